// capi.cu — the extern "C" boundary declared in include/rbfmap.h.
#include "mapping.cuh"
#include "nccl_dyn.cuh"
#include "pum.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <limits>
#include <mutex>
#include <vector>

namespace rbfmap {

// ---------------------------------------------------------------------------------------------
// basis-function parameters (host): same derivation and argument checks as the reference functor
// constructors (mapping/impl/BasisFunctions.hpp:120-124, 159-166, 229-252, 307-316, 361-370, ...)
// ---------------------------------------------------------------------------------------------
namespace {
double hostGaussian(double radius, double shape, double support, double deltaY)
{
  if (radius > support)
    return 0.0;
  const double sr = shape * radius;
  return std::exp(-(sr * sr)) - deltaY;
}
const char *compactName(int kind)
{
  switch (kind) {
  case RBFMAP_COMPACT_TPS_C2: return "compact thin-plate-splines c2";
  case RBFMAP_COMPACT_POLYNOMIAL_C0: return "compact polynomial c0";
  case RBFMAP_COMPACT_POLYNOMIAL_C2: return "compact polynomial c2";
  case RBFMAP_COMPACT_POLYNOMIAL_C4: return "compact polynomial c4";
  default: return "compact polynomial c6"; // the reference's C8 constructor also says "c6" (BasisFunctions.hpp:569)
  }
}
} // namespace

RbfParams makeRbfParams(int kind, double supportRadius, double shapeParameter)
{
  RbfParams f{kind, 0.0, 0.0, 0.0};
  switch (kind) {
  case RBFMAP_THIN_PLATE_SPLINES:
  case RBFMAP_VOLUME_SPLINES:
    break;
  case RBFMAP_MULTIQUADRICS:
    f.p1 = shapeParameter * shapeParameter;
    break;
  case RBFMAP_INVERSE_MULTIQUADRICS:
    if (!(shapeParameter > 0.0))
      throw MapError(RBFMAP_ERR_INVALID, "Shape parameter for radial-basis-function inverse multiquadric has to be larger than zero. Please update the \"shape-parameter\" attribute.");
    f.p1 = shapeParameter * shapeParameter;
    break;
  case RBFMAP_GAUSSIAN: {
    const double support = supportRadius > 0.0 ? supportRadius : std::numeric_limits<double>::infinity();
    if (!(shapeParameter > 0.0))
      throw MapError(RBFMAP_ERR_INVALID, "Shape parameter for radial-basis-function gaussian has to be larger than zero. Please update the \"shape-parameter\" attribute.");
    const double threshold = std::sqrt(-std::log(1e-9)) / shapeParameter; // cutoffThreshold, BasisFunctions.hpp:281
    f.p1                   = shapeParameter;
    f.p2                   = std::min(support, threshold);
    f.p3                   = 0.0;
    if (support < std::numeric_limits<double>::infinity())
      f.p3 = hostGaussian(support, f.p1, f.p2, 0.0);
    break;
  }
  case RBFMAP_COMPACT_TPS_C2:
  case RBFMAP_COMPACT_POLYNOMIAL_C0:
  case RBFMAP_COMPACT_POLYNOMIAL_C2:
  case RBFMAP_COMPACT_POLYNOMIAL_C4:
  case RBFMAP_COMPACT_POLYNOMIAL_C6:
  case RBFMAP_COMPACT_POLYNOMIAL_C8:
    if (!(supportRadius > 0.0))
      throw MapError(RBFMAP_ERR_INVALID, std::string("Support radius for radial-basis-function ") + compactName(kind) +
                                             " has to be larger than zero. Please update the \"support-radius\" attribute.");
    f.p1 = 1.0 / supportRadius;
    break;
  default:
    throw MapError(RBFMAP_ERR_INVALID, "unknown basis function");
  }
  return f;
}

double basisSupportRadius(const RbfParams &f)
{
  if (!basisHasCompactSupport(f.kind))
    return std::numeric_limits<double>::max();
  if (f.kind == RBFMAP_GAUSSIAN)
    return f.p2;
  return 1.0 / f.p1;
}

// ---------------------------------------------------------------------------------------------
DeviceMapping::DeviceMapping(const rbfmap_config &cfg)
    : _cfg(cfg)
{
  stats = rbfmap_stats{};
  if (cfg.dimensions != 2 && cfg.dimensions != 3)
    throw MapError(RBFMAP_ERR_INVALID, "dimensions has to be 2 or 3");
  if (cfg.constraint < RBFMAP_CONSISTENT || cfg.constraint > RBFMAP_SCALED_CONSISTENT_VOLUME)
    throw MapError(RBFMAP_ERR_INVALID, "unknown mapping constraint");
  _rbf = makeRbfParams(cfg.basis, cfg.support_radius, cfg.shape_parameter);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    throw MapError(RBFMAP_ERR_CUDA, "rbfmap: no CUDA device available; this backend has no CPU fallback.");
  }
  if (cfg.execution_mode != RBFMAP_MINIMAL_MEMORY && cfg.execution_mode != RBFMAP_MINIMAL_COMPUTE)
    throw MapError(RBFMAP_ERR_INVALID, "unknown execution mode");
  if (cfg.shard_mode != RBFMAP_SHARD_DUPLICATE && cfg.shard_mode != RBFMAP_SHARD_EXCHANGE)
    throw MapError(RBFMAP_ERR_INVALID, "unknown shard mode");
  if (cfg.device_id >= 0) {
    if (cfg.device_id >= count)
      throw MapError(RBFMAP_ERR_CUDA, "rbfmap: gpu-device-id out of range.");
    _device = cfg.device_id;
  } else {
    RBF_CUDA(cudaGetDevice(&_device));
  }
  // the creating thread's current device is left untouched: every entry point selects the mapping's device itself
  DeviceScope scope(_device);
  RBF_CUDA(cudaStreamCreateWithFlags(&_stream, cudaStreamNonBlocking));
  configureMemoryPool();
}

DeviceMapping::~DeviceMapping()
{
  if (_stream) {
    DeviceScope scope(_device);
    cudaStreamDestroy(_stream);
  }
}

void DeviceMapping::uploadReplicated(double *dDst, const double *hSrc, int64_t count, cudaStream_t s) const
{
  if (count <= 0)
    return;
  if (!_group.connected()) {
    RBF_CUDA(cudaMemcpyAsync(dDst, hSrc, static_cast<size_t>(count) * sizeof(double), cudaMemcpyHostToDevice, s));
    return;
  }
  const int64_t slice = _group.sliceOf(count);
  const int64_t b = std::min(count, slice * _group.rank), e = std::min(count, b + slice);
  if (e > b)
    RBF_CUDA(cudaMemcpyAsync(dDst + b, hSrc + b, static_cast<size_t>(e - b) * sizeof(double), cudaMemcpyHostToDevice, s));
  _group.allGatherInPlace(dDst, slice, s);
}

namespace {
__global__ void tagBoxKernel(const double *__restrict__ xyz, long long n, int dim, double lx, double ly, double lz, double hx, double hy, double hz,
                             unsigned char *__restrict__ tags)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  // mesh/BoundingBox.cpp:70-81: outside iff coords[d] < min[d] || coords[d] > max[d]
  bool inside = !(x < lx || x > hx) && !(y < ly || y > hy);
  if (dim == 3)
    inside = inside && !(z < lz || z > hz);
  if (inside)
    tags[i] = 1;
}
__global__ void ownedBoundsKernel(const double *__restrict__ xyz, long long n, const unsigned char *__restrict__ owner, unsigned long long *__restrict__ keys)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n || !owner[i])
    return;
  for (int d = 0; d < 3; ++d) {
    atomicMin(&keys[d], encodeOrdered(xyz[3 * i + d]));
    atomicMax(&keys[3 + d], encodeOrdered(xyz[3 * i + d]));
  }
}
} // namespace

void DeviceMapping::tagInsideBox(const double *dXyz, int64_t n, const double lo[3], const double hi[3], double margin, unsigned char *hTags)
{
  if (n <= 0)
    return;
  DevBuf<unsigned char> dTags;
  dTags.alloc(n);
  RBF_CUDA(cudaMemcpyAsync(dTags.p, hTags, n, cudaMemcpyHostToDevice, _stream));
  // BoundingBox::expandBy (mesh/BoundingBox.cpp:154-160): min - value, max + value
  tagBoxKernel<<<divUp(n, 256), 256, 0, _stream>>>(dXyz, n, _cfg.dimensions, lo[0] - margin, lo[1] - margin, lo[2] - margin, hi[0] + margin, hi[1] + margin,
                                                  hi[2] + margin, dTags.p);
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
  RBF_CUDA(cudaMemcpyAsync(hTags, dTags.p, n, cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
}

// RadialBasisFctBaseMapping.hpp:124-157
void DeviceMapping::tagMeshFirstRound(const double *hFilterXyz, int64_t nFilter, const double *hOtherXyz, int64_t nOther, unsigned char *hTags)
{
  if (nOther == 0 || nFilter == 0)
    return; // :137-138 ranks not at the interface never hold interface vertices
  if (!basisHasCompactSupport(_rbf.kind)) {
    std::memset(hTags, 1, static_cast<size_t>(nFilter)); // :154 tagAll
    return;
  }
  AllocStreamScope allocScope(_stream);
  DevBuf<double>   dFilter, dOther;
  dFilter.alloc(3 * nFilter);
  dOther.alloc(3 * nOther);
  RBF_CUDA(cudaMemcpyAsync(dFilter.p, hFilterXyz, 3 * nFilter * sizeof(double), cudaMemcpyHostToDevice, _stream));
  RBF_CUDA(cudaMemcpyAsync(dOther.p, hOtherXyz, 3 * nOther * sizeof(double), cudaMemcpyHostToDevice, _stream));
  double lo[3], hi[3];
  computeBounds(dOther.p, nOther, lo, hi, _stream);
  tagInsideBox(dFilter.p, nFilter, lo, hi, basisSupportRadius(_rbf), hTags); // :141-151
}

// RadialBasisFctBaseMapping.hpp:163-190
void DeviceMapping::tagMeshSecondRound(const double *hFilterXyz, int64_t nFilter, const unsigned char *hOwner, unsigned char *hTags)
{
  if (!basisHasCompactSupport(_rbf.kind) || nFilter == 0)
    return; // :167-168 tags are not changed
  AllocStreamScope           allocScope(_stream);
  DevBuf<double>             dFilter;
  DevBuf<unsigned char>      dOwner;
  DevBuf<unsigned long long> keys;
  dFilter.alloc(3 * nFilter);
  dOwner.alloc(nFilter);
  keys.alloc(6);
  unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
  RBF_CUDA(cudaMemcpyAsync(dFilter.p, hFilterXyz, 3 * nFilter * sizeof(double), cudaMemcpyHostToDevice, _stream));
  RBF_CUDA(cudaMemcpyAsync(dOwner.p, hOwner, nFilter, cudaMemcpyHostToDevice, _stream));
  RBF_CUDA(cudaMemcpyAsync(keys.p, init, sizeof(init), cudaMemcpyHostToDevice, _stream));
  ownedBoundsKernel<<<divUp(nFilter, 256), 256, 0, _stream>>>(dFilter.p, nFilter, dOwner.p, keys.p); // :180-186 box around all owned vertices
  RBF_LAUNCHED();
  unsigned long long out[6];
  RBF_CUDA(cudaMemcpyAsync(out, keys.p, sizeof(out), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  if (out[0] == ~0ull)
    return; // no owned vertex: the default box is not expanded (BoundingBox::expandBy), nothing lies inside
  double lo[3], hi[3];
  for (int d = 0; d < 3; ++d) {
    lo[d] = decodeOrdered(out[d]);
    hi[d] = decodeOrdered(out[3 + d]);
  }
  tagInsideBox(dFilter.p, nFilter, lo, hi, basisSupportRadius(_rbf), hTags); // :187-189
}

double PhaseTrace::now()
{
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
PhaseTrace::PhaseTrace(const char *what_, cudaStream_t stream)
    : s(stream), what(what_)
{
  const char *e = getenv("RBFMAP_TRACE");
  on            = e && e[0] == '1';
  if (on) {
    cudaStreamSynchronize(s);
    t0 = now();
  }
}
void PhaseTrace::phase(const char *name)
{
  if (!on)
    return;
  cudaStreamSynchronize(s);
  const double t = now();
  fprintf(stderr, "[rbfmap trace] %s: %-28s %9.3f ms\n", what, name, t - t0);
  t0 = t;
}

bool pumInspect(DeviceMapping *m, int64_t &nClusters, const double *&centers, const PumMembership *&mem);
void debugDistSqrt(const double *dX, long long n, double *dOut);
void debugPairBasis(const RbfParams &f, bool dim3, const double *dA, const double *dB, long long n, int batched, double *dOut);

namespace {
__global__ void evalBasisKernel(RbfParams f, const double *__restrict__ r, long long n, double *__restrict__ out)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n)
    out[i] = evalBasis(r[i], f);
}
thread_local std::string g_createError;
} // namespace

} // namespace rbfmap

using namespace rbfmap;

struct rbfmap_handle {
  std::unique_ptr<DeviceMapping> impl;
  std::string                    lastError;
  std::string                    name;
  // staging buffers of the host-pointer API
  DevBuf<double> dInXyz, dOutXyz, dIn, dOut;
  // scaled-consistent constraints (Mapping.cpp:175-246): connectivity of Mapping::input() [0] and output() [1] as the caller
  // handed it over, and the per-vertex integration weights derived from it at computeMapping
  struct Connectivity {
    std::vector<int32_t> elements;
    int                  verticesPerElement = 0;
    DevBuf<double>       weights; // integral of the nodal hat function of every vertex: sum over its elements of measure / k
    int64_t              nVertices = 0;
  } conn[2];
};

namespace {
template <typename F>
int guarded(rbfmap_handle *h, F &&f)
{
  try {
    // the mapping lives on one device (gpu-device-id); select it for the duration of the call, whichever thread calls
    // and whatever device that thread had current, and restore the caller's device afterwards
    DeviceScope scope(h && h->impl ? h->impl->device() : -1);
    f();
    if (h)
      h->lastError.clear();
    return RBFMAP_OK;
  } catch (const MapError &e) {
    (h ? h->lastError : g_createError) = e.what();
    cudaGetLastError();
    return e.code;
  } catch (const std::exception &e) {
    (h ? h->lastError : g_createError) = e.what();
    return RBFMAP_ERR_INVALID;
  }
}
} // namespace

// ---- scaled-consistent mapping (Mapping::scaleConsistentMapping, Mapping.cpp:175-246; mesh::integrateSurface / integrateVolume,
// mesh/Utils.cpp:11-70; element measures math/geometry.cpp:93-123, mesh/Edge.cpp:23-27) ----
namespace rbfmap {
namespace scaled {
__global__ void elementWeightsKernel(const double *__restrict__ xyz, const int *__restrict__ el, long long nEl, int k, int dim, double *__restrict__ w)
{
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e >= nEl)
    return;
  double p[4][3];
  for (int a = 0; a < k; ++a)
    for (int d = 0; d < 3; ++d)
      p[a][d] = (d < dim) ? xyz[3 * static_cast<long long>(el[k * e + a]) + d] : 0.0;
  double measure;
  if (k == 2) { // Edge::getLength
    const double dx = p[1][0] - p[0][0], dy = p[1][1] - p[0][1], dz = p[1][2] - p[0][2];
    measure         = sqrt(dx * dx + dy * dy + dz * dz);
  } else if (k == 3) { // triangleArea: 0.5 |A x B|
    const double ax = p[1][0] - p[0][0], ay = p[1][1] - p[0][1], az = p[1][2] - p[0][2];
    const double bx = p[2][0] - p[0][0], by = p[2][1] - p[0][1], bz = p[2][2] - p[0][2];
    if (dim == 2) {
      measure = 0.5 * fabs(ax * by - ay * bx);
    } else {
      const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
      measure         = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
    }
  } else { // tetraVolume: |(a - d) . ((b - d) x (c - d))| / 6
    double u[3], v[3], t[3];
    for (int d = 0; d < 3; ++d) {
      u[d] = p[0][d] - p[3][d];
      v[d] = p[1][d] - p[3][d];
      t[d] = p[2][d] - p[3][d];
    }
    const double cx = v[1] * t[2] - v[2] * t[1], cy = v[2] * t[0] - v[0] * t[2], cz = v[0] * t[1] - v[1] * t[0];
    measure         = fabs(u[0] * cx + u[1] * cy + u[2] * cz) / 6.0;
  }
  const double share = measure / static_cast<double>(k); // 0.5 length, area / 3, volume / 4 per vertex (mesh/Utils.cpp:24, 34, 65)
  for (int a = 0; a < k; ++a)
    atomicAdd(&w[el[k * e + a]], share);
}

// integral[c] += sum_v w[v] values[v * dims + c]
__global__ void weightedSumKernel(const double *__restrict__ w, const double *__restrict__ values, long long n, int dims, double *__restrict__ integral)
{
  for (int c = 0; c < dims; ++c) {
    double s = 0.0;
    for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < n; v += static_cast<long long>(gridDim.x) * blockDim.x)
      s += w[v] * values[v * dims + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s != 0.0)
      atomicAdd(&integral[c], s);
  }
}

// Mapping.cpp:230-235: factor = integralInput / integralOutput, 1 if the output integral is zero
__global__ void scaleKernel(double *__restrict__ out, long long n, int dims, const double *__restrict__ integrals /* dims input, dims output */)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n * dims)
    return;
  const int    c   = static_cast<int>(i % dims);
  const double den = integrals[dims + c];
  out[i] *= (den == 0.0) ? 1.0 : integrals[c] / den;
}
} // namespace scaled
} // namespace rbfmap
using namespace rbfmap::scaled;

namespace rbfmap {
namespace scaled {
// combined[v * D + off + c] <-> sample[v * dims + c]
__global__ void interleaveKernel(const double *__restrict__ sample, long long n, int dims, int D, int off, double *__restrict__ combined)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n * dims)
    combined[(i / dims) * D + off + (i % dims)] = sample[i];
}
__global__ void deinterleaveKernel(const double *__restrict__ combined, long long n, int dims, int D, int off, double *__restrict__ sample)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n * dims)
    sample[i] = combined[(i / dims) * D + off + (i % dims)];
}
} // namespace scaled
} // namespace rbfmap

static bool isScaledConsistent(const DeviceMapping *m)
{
  const int c = m->config().constraint;
  return c == RBFMAP_SCALED_CONSISTENT_SURFACE || c == RBFMAP_SCALED_CONSISTENT_VOLUME;
}

// vertices per element that Mapping.cpp:190-192 demands: 2-D surface -> edges, 2-D volume / 3-D surface -> triangles, 3-D volume -> tetrahedra
static int requiredElement(const DeviceMapping *m)
{
  const bool volume = m->config().constraint == RBFMAP_SCALED_CONSISTENT_VOLUME;
  return m->config().dimensions == 2 ? (volume ? 3 : 2) : (volume ? 4 : 3);
}

// called after a successful computeMapping with the device copies of Mapping::input() / output()
static void computeIntegrationWeights(rbfmap_handle *h, const double *dInputXyz, int64_t nInput, const double *dOutputXyz, int64_t nOutput)
{
  DeviceMapping *m = h->impl.get();
  if (!isScaledConsistent(m))
    return;
  cudaStream_t     s = m->stream();
  AllocStreamScope allocScope(s);
  const int        k = requiredElement(m);
  const double    *xyz[2] = {dInputXyz, dOutputXyz};
  const int64_t    nv[2]  = {nInput, nOutput};
  for (int which = 0; which < 2; ++which) {
    auto &cn     = h->conn[which];
    cn.nVertices = nv[which];
    cn.weights.release();
    if (nv[which] == 0 || cn.verticesPerElement != k || cn.elements.empty())
      continue; // reported by map() with the reference's message
    const int64_t nEl = static_cast<int64_t>(cn.elements.size()) / k;
    for (int32_t v : cn.elements)
      if (v < 0 || v >= nv[which])
        throw MapError(RBFMAP_ERR_INVALID, "connectivity refers to a vertex that does not exist");
    DevBuf<int> dEl;
    dEl.alloc(cn.elements.size());
    cn.weights.alloc(nv[which]);
    RBF_CUDA(cudaMemcpyAsync(dEl.p, cn.elements.data(), cn.elements.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    RBF_CUDA(cudaMemsetAsync(cn.weights.p, 0, nv[which] * sizeof(double), s));
    elementWeightsKernel<<<divUp(nEl, 256), 256, 0, s>>>(xyz[which], dEl.p, nEl, k, m->config().dimensions, cn.weights.p);
    RBF_CUDA(cudaGetLastError());
    RBF_CUDA(cudaStreamSynchronize(s)); // dEl is released on return
  }
}

static void scaleConsistentMapping(rbfmap_handle *h, const double *dIn, int dims, double *dOut)
{
  DeviceMapping *m = h->impl.get();
  const int64_t  nIn = m->nInput(), nOut = m->nOutput();
  if (nIn == 0 || nOut == 0)
    return; // Mapping.cpp:179-181
  if (m->group().active())
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "Scaled-consistent mapping is only supported for serial participants."); // :186-187
  static const char *kindName[5] = {"", "", "Edges", "Triangles", "Tetrahedra"};
  const int          k           = requiredElement(m);
  for (int which = 0; which < 2; ++which)
    if (!h->conn[which].weights.p || h->conn[which].nVertices != (which == 0 ? nIn : nOut)) // :194-208
      throw MapError(RBFMAP_ERR_INVALID, std::string(kindName[k]) + " connectivity information is missing for the mesh \"" + (which == 0 ? "input" : "output") +
                                             "\". Scaled consistent mapping requires connectivity information.");
  cudaStream_t     s = m->stream();
  AllocStreamScope allocScope(s);
  DevBuf<double>   integrals;
  integrals.alloc(2 * static_cast<size_t>(dims));
  RBF_CUDA(cudaMemsetAsync(integrals.p, 0, 2 * dims * sizeof(double), s));
  const int grid = std::min(1024, std::max(1, divUp(std::max(nIn, nOut), 256)));
  weightedSumKernel<<<grid, 256, 0, s>>>(h->conn[0].weights.p, dIn, nIn, dims, integrals.p);
  weightedSumKernel<<<grid, 256, 0, s>>>(h->conn[1].weights.p, dOut, nOut, dims, integrals.p + dims);
  scaleKernel<<<divUp(nOut * dims, 256), 256, 0, s>>>(dOut, nOut, dims, integrals.p);
  RBF_CUDA(cudaGetLastError());
  RBF_CUDA(cudaStreamSynchronize(s));
}

extern "C" {

int rbfmap_set_connectivity(rbfmap_handle *h, int mesh, const int32_t *elements, int64_t n_elements, int vertices_per_element)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    if ((mesh != 0 && mesh != 1) || n_elements < 0 || vertices_per_element < 2 || vertices_per_element > 4 || (n_elements > 0 && !elements))
      throw MapError(RBFMAP_ERR_INVALID, "invalid connectivity");
    if (h->impl->hasComputedMapping())
      throw MapError(RBFMAP_ERR_STATE, "Set the mesh connectivity before computeMapping() (the integration weights belong to the computed mapping).");
    auto &cn              = h->conn[mesh];
    cn.verticesPerElement = vertices_per_element;
    cn.elements.assign(elements, elements + n_elements * vertices_per_element);
    cn.weights.release();
  });
}

void rbfmap_default_config(rbfmap_config *cfg)
{
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->variant              = RBFMAP_PUM_DIRECT;
  cfg->constraint           = RBFMAP_CONSISTENT;
  cfg->dimensions           = 3;
  cfg->basis                = RBFMAP_COMPACT_POLYNOMIAL_C6;
  cfg->polynomial           = RBFMAP_POLYNOMIAL_SEPARATE; // MappingConfiguration.cpp:227
  cfg->vertices_per_cluster = 50;                         // :243-248
  cfg->relative_overlap     = 0.15;
  cfg->project_to_input     = 1;
  cfg->solver_rtol          = 1e-9; // :237
  cfg->max_iterations       = 1000000;
  cfg->device_id            = -1;
  cfg->execution_mode       = RBFMAP_MINIMAL_MEMORY; // PartitionOfUnityMapping.hpp:57 (computeEvaluationOffline = false)
  cfg->shard_mode           = RBFMAP_SHARD_DUPLICATE;
  cfg->deterministic        = 0;
  cfg->rescaling            = 1;    // PetRadialBasisFctMapping.hpp:126-127 (useRescaling = true)
  cfg->output_filter        = 1e-9; // PetRadialBasisFctMapping.hpp:675, 799
}

int rbfmap_create(const rbfmap_config *cfg, rbfmap_handle **out)
{
  if (!cfg || !out)
    return RBFMAP_ERR_INVALID;
  *out = nullptr;
  return guarded(nullptr, [&] {
    auto h = std::make_unique<rbfmap_handle>();
    switch (cfg->variant) {
    case RBFMAP_PUM_DIRECT: h->impl = makePumMapping(*cfg); break;
    case RBFMAP_GLOBAL_DIRECT: h->impl = makeGlobalDirectMapping(*cfg); break;
    case RBFMAP_GLOBAL_ITERATIVE: h->impl = makeGlobalIterativeMapping(*cfg); break;
    default: throw MapError(RBFMAP_ERR_INVALID, "unknown mapping variant");
    }
    h->name = h->impl->name();
    *out    = h.release();
  });
}

int rbfmap_compute_mapping_device(rbfmap_handle *h, const double *d_input_xyz, int64_t n_input, const double *d_output_xyz, int64_t n_output)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    if (n_input < 0 || n_output < 0)
      throw MapError(RBFMAP_ERR_INVALID, "negative mesh size");
    h->impl->computeMapping(d_input_xyz, n_input, d_output_xyz, n_output);
    computeIntegrationWeights(h, d_input_xyz, n_input, d_output_xyz, n_output);
  });
}

int rbfmap_compute_mapping(rbfmap_handle *h, const double *input_xyz, int64_t n_input, const double *output_xyz, int64_t n_output)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    if (n_input < 0 || n_output < 0)
      throw MapError(RBFMAP_ERR_INVALID, "negative mesh size");
    cudaStream_t     s = h->impl->stream();
    AllocStreamScope allocScope(s);
    if (h->impl->computeMappingHost(input_xyz, n_input, output_xyz, n_output)) {
      if (isScaledConsistent(h->impl.get())) { // the integration weights need the coordinates once more (small serial meshes)
        h->dInXyz.alloc(3 * std::max<int64_t>(n_input, 1));
        h->dOutXyz.alloc(3 * std::max<int64_t>(n_output, 1));
        if (n_input > 0)
          RBF_CUDA(cudaMemcpyAsync(h->dInXyz.p, input_xyz, 3 * n_input * sizeof(double), cudaMemcpyHostToDevice, s));
        if (n_output > 0)
          RBF_CUDA(cudaMemcpyAsync(h->dOutXyz.p, output_xyz, 3 * n_output * sizeof(double), cudaMemcpyHostToDevice, s));
        computeIntegrationWeights(h, h->dInXyz.p, n_input, h->dOutXyz.p, n_output);
        h->dInXyz.release();
        h->dOutXyz.release();
      }
      return;
    }
    const ShardGroup &grp = h->impl->group();
    h->dInXyz.alloc(std::max<int64_t>(grp.padded(3 * n_input), 1));
    h->dOutXyz.alloc(std::max<int64_t>(grp.padded(3 * n_output), 1));
    h->impl->uploadReplicated(h->dInXyz.p, input_xyz, 3 * n_input, s);
    h->impl->uploadReplicated(h->dOutXyz.p, output_xyz, 3 * n_output, s);
    RBF_CUDA(cudaStreamSynchronize(s));
    h->impl->computeMapping(h->dInXyz.p, n_input, h->dOutXyz.p, n_output);
    computeIntegrationWeights(h, h->dInXyz.p, n_input, h->dOutXyz.p, n_output);
    // the mapping keeps its own copies; release the staging buffers
    h->dInXyz.release();
    h->dOutXyz.release();
  });
}

struct rbfmap_cache {
  std::unique_ptr<DeviceMapping::JitCache> impl;
  int                                      device = -1;
};

extern "C" {
int rbfmap_compute_mapping_just_in_time(rbfmap_handle *h, const double *mesh_xyz, int64_t n)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    if (n < 0 || (n > 0 && !mesh_xyz))
      throw MapError(RBFMAP_ERR_INVALID, "invalid mesh");
    h->impl->jitComputeMapping(mesh_xyz, n);
  });
}
int rbfmap_cache_create(rbfmap_handle *h, int dims, rbfmap_cache **out)
{
  if (!h || !out)
    return RBFMAP_ERR_INVALID;
  *out = nullptr;
  return guarded(h, [&] {
    auto c  = std::make_unique<rbfmap_cache>();
    c->impl   = h->impl->jitCreateCache(dims);
    c->device = h->impl->device();
    *out      = c.release();
  });
}
int rbfmap_cache_reset(rbfmap_handle *h, rbfmap_cache *cache)
{
  if (!h || !cache)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] { h->impl->jitResetCache(*cache->impl); });
}
void rbfmap_cache_destroy(rbfmap_cache *cache)
{
  if (!cache)
    return;
  DeviceScope scope(cache->device);
  delete cache;
}
int  rbfmap_cache_update(rbfmap_handle *h, rbfmap_cache *cache, const double *values)
{
  if (!h || !cache)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] { h->impl->jitUpdateCache(*cache->impl, values); });
}
int rbfmap_map_consistent_at(rbfmap_handle *h, rbfmap_cache *cache, const double *coords, int64_t n_points, double *out)
{
  if (!h || !cache)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] { h->impl->jitMapConsistentAt(*cache->impl, coords, n_points, out); });
}
int rbfmap_map_conservative_at(rbfmap_handle *h, rbfmap_cache *cache, const double *coords, int64_t n_points, const double *values)
{
  if (!h || !cache)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] { h->impl->jitMapConservativeAt(*cache->impl, coords, n_points, values); });
}
int rbfmap_complete_just_in_time_mapping(rbfmap_handle *h, rbfmap_cache *cache, double *out)
{
  if (!h || !cache)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] { h->impl->jitComplete(*cache->impl, out); });
}
} // extern "C"

static void mapDeviceImpl(rbfmap_handle *h, int mode /*0 dispatch, 1 consistent, 2 conservative*/, const double *d_in, int dims, double *d_out)
{
  DeviceMapping *m    = h->impl.get();
  if (dims < 1)
    throw MapError(RBFMAP_ERR_INVALID, "data dimension has to be positive");
  // The reference reaches the two virtuals only through Mapping::map(), which dispatches on the constraint
  // (Mapping.cpp:158-173): the mapping was computed for one direction (the meshes are swapped for conservative
  // constraints), so forcing the other one would read and write buffers of the wrong size.
  if ((mode == 2 && !m->isConservative()) || (mode == 1 && m->isConservative()))
    throw MapError(RBFMAP_ERR_STATE, std::string("The mapping was computed with a ") + (m->isConservative() ? "conservative" : "consistent") +
                                         " constraint; " + (mode == 2 ? "mapConservative" : "mapConsistent") + " is not available on it.");
  const bool     cons = mode == 2 || (mode == 0 && m->isConservative());
  if (cons)
    m->mapConservative(d_in, dims, d_out);
  else
    m->mapConsistent(d_in, dims, d_out);
  RBF_CUDA(cudaStreamSynchronize(m->stream()));
  // Mapping.cpp:167-169: scaled-consistent = mapConsistent + scaleConsistentMapping (only through the dispatching map())
  if (mode == 0 && isScaledConsistent(m))
    scaleConsistentMapping(h, d_in, dims, d_out);
}

static int mapHostImpl(rbfmap_handle *h, int mode, const double *in, int dims, double *out, double *guess = nullptr)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    DeviceMapping *m = h->impl.get();
    if (!m->hasComputedMapping())
      throw MapError(RBFMAP_ERR_STATE, "The mapping has not been computed; call computeMapping() first.");
    if (dims < 1)
      throw MapError(RBFMAP_ERR_INVALID, "data dimension has to be positive");
    cudaStream_t     s = m->stream();
    AllocStreamScope allocScope(s);
    const int64_t    nIn  = m->nInput() * dims;
    const int64_t nOut = m->nOutput() * dims;
    const int64_t nInPad = std::max<int64_t>(m->group().padded(nIn), 1);
    if (h->dIn.n < static_cast<size_t>(nInPad))
      h->dIn.alloc(nInPad);
    if (h->dOut.n < static_cast<size_t>(std::max<int64_t>(nOut, 1)))
      h->dOut.alloc(std::max<int64_t>(nOut, 1));
    m->uploadReplicated(h->dIn.p, in, nIn, s); // distributed: 1/world of the values per rank over PCIe, all-gather over NVLink
    if (guess)
      m->loadInitialGuess(guess, dims);
    mapDeviceImpl(h, mode, h->dIn.p, dims, h->dOut.p);
    if (guess)
      m->storeInitialGuess(guess, dims);
    if (nOut > 0)
      RBF_CUDA(cudaMemcpyAsync(out, h->dOut.p, nOut * sizeof(double), cudaMemcpyDeviceToHost, s));
    RBF_CUDA(cudaStreamSynchronize(s));
  });
}

int rbfmap_map_sliced(rbfmap_handle *h, const double *in, int dims, double *out, int64_t *begin, int64_t *end)
{
  if (!h || !begin || !end)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    DeviceMapping *m = h->impl.get();
    if (!m->hasComputedMapping())
      throw MapError(RBFMAP_ERR_STATE, "The mapping has not been computed; call computeMapping() first.");
    if (dims < 1)
      throw MapError(RBFMAP_ERR_INVALID, "data dimension has to be positive");
    const ShardGroup &g = m->group();
    if (g.active() && !g.connected() && m->mapResultIsPartial())
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "a group without communicator cannot complete the result slices; use rbfmap_map.");
    cudaStream_t     s = m->stream();
    AllocStreamScope allocScope(s);
    const int64_t    nIn = m->nInput() * dims, nOutV = m->nOutput();
    const int64_t    sliceV = g.connected() ? g.sliceOf(nOutV) : nOutV; // output vertices per rank
    const int64_t    b = std::min(nOutV, sliceV * (g.connected() ? g.rank : 0)), e = std::min(nOutV, b + sliceV);
    const int64_t    nOut = nOutV * dims, nOutPad = (g.connected() ? sliceV * g.world : nOutV) * dims;
    const int64_t    nInPad = std::max<int64_t>(g.padded(nIn), 1);
    if (h->dIn.n < static_cast<size_t>(nInPad))
      h->dIn.alloc(nInPad);
    if (h->dOut.n < static_cast<size_t>(std::max<int64_t>(nOutPad, 1)))
      h->dOut.alloc(std::max<int64_t>(nOutPad, 1));
    m->uploadReplicated(h->dIn.p, in, nIn, s);
    mapDeviceImpl(h, 0, h->dIn.p, dims, h->dOut.p);
    if (g.connected() && m->mapResultIsPartial()) { // every rank contributes its owned vertices; the slices are summed over NVLink
      if (nOutPad > nOut)
        RBF_CUDA(cudaMemsetAsync(h->dOut.p + nOut, 0, static_cast<size_t>(nOutPad - nOut) * sizeof(double), s));
      g.reduceScatterSumInPlace(h->dOut.p, sliceV * dims, s);
    }
    if (e > b)
      RBF_CUDA(cudaMemcpyAsync(out + b * dims, h->dOut.p + b * dims, static_cast<size_t>(e - b) * dims * sizeof(double), cudaMemcpyDeviceToHost, s));
    RBF_CUDA(cudaStreamSynchronize(s));
    *begin = b;
    *end   = e;
  });
}

int rbfmap_map(rbfmap_handle *h, const double *in, int dims, double *out) { return mapHostImpl(h, 0, in, dims, out); }
int rbfmap_map_consistent(rbfmap_handle *h, const double *in, int dims, double *out) { return mapHostImpl(h, 1, in, dims, out); }
int rbfmap_map_conservative(rbfmap_handle *h, const double *in, int dims, double *out) { return mapHostImpl(h, 2, in, dims, out); }

int rbfmap_map_batch(rbfmap_handle *h, int count, const double *const *in, const int *dims, double *const *out)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    DeviceMapping *m = h->impl.get();
    if (!m->hasComputedMapping())
      throw MapError(RBFMAP_ERR_STATE, "The mapping has not been computed; call computeMapping() first.");
    if (count < 0 || (count > 0 && (!in || !dims || !out)))
      throw MapError(RBFMAP_ERR_INVALID, "invalid batch");
    int D = 0;
    for (int k = 0; k < count; ++k) {
      if (dims[k] < 1 || !in[k] || !out[k])
        throw MapError(RBFMAP_ERR_INVALID, "invalid sample in the batch");
      D += dims[k];
    }
    if (count == 0)
      return;
    cudaStream_t     s = m->stream();
    AllocStreamScope allocScope(s);
    const int64_t    nIn = m->nInput(), nOut = m->nOutput();
    int              maxDims = 0;
    for (int k = 0; k < count; ++k)
      maxDims = std::max(maxDims, dims[k]);
    DevBuf<double> dAllIn, dAllOut, dStage;
    dAllIn.alloc(std::max<int64_t>(nIn * D, 1));
    dAllOut.alloc(std::max<int64_t>(nOut * D, 1));
    dStage.alloc(std::max<int64_t>(m->group().padded(std::max(nIn, nOut) * maxDims), 1));
    // all samples side by side as the components of one field: the per-cluster kernels gather, factor-load and evaluate the
    // basis function once per group of three components instead of once per sample
    for (int k = 0, off = 0; k < count; off += dims[k], ++k) {
      m->uploadReplicated(dStage.p, in[k], nIn * dims[k], s);
      if (nIn > 0)
        interleaveKernel<<<divUp(nIn * dims[k], 256), 256, 0, s>>>(dStage.p, nIn, dims[k], D, off, dAllIn.p);
    }
    RBF_CUDA(cudaGetLastError());
    mapDeviceImpl(h, 0, dAllIn.p, D, dAllOut.p);
    for (int k = 0, off = 0; k < count; off += dims[k], ++k) {
      if (nOut > 0) {
        deinterleaveKernel<<<divUp(nOut * dims[k], 256), 256, 0, s>>>(dAllOut.p, nOut, dims[k], D, off, dStage.p);
        RBF_CUDA(cudaMemcpyAsync(out[k], dStage.p, static_cast<size_t>(nOut) * dims[k] * sizeof(double), cudaMemcpyDeviceToHost, s));
      }
    }
    RBF_CUDA(cudaGetLastError());
    RBF_CUDA(cudaStreamSynchronize(s));
  });
}

int64_t rbfmap_initial_guess_size(const rbfmap_handle *h) { return h ? h->impl->initialGuessSize() : 0; }

int rbfmap_map_with_initial_guess(rbfmap_handle *h, const double *in, int dims, double *out, double *guess)
{
  return mapHostImpl(h, 0, in, dims, out, (h && h->impl->initialGuessSize() > 0) ? guess : nullptr);
}

int rbfmap_map_device(rbfmap_handle *h, const double *d_in, int dims, double *d_out)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    if (!h->impl->hasComputedMapping())
      throw MapError(RBFMAP_ERR_STATE, "The mapping has not been computed; call computeMapping() first.");
    mapDeviceImpl(h, 0, d_in, dims, d_out);
  });
}

int rbfmap_tag_mesh_first_round(rbfmap_handle *h, const double *input_xyz, int64_t n_input, const double *output_xyz, int64_t n_output, unsigned char *tags)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    if (n_input < 0 || n_output < 0 || !tags)
      throw MapError(RBFMAP_ERR_INVALID, "invalid mesh / tag buffer");
    const bool cons = h->impl->isConservative(); // the remote (filter) mesh is output() for conservative constraints
    h->impl->tagMeshFirstRound(cons ? output_xyz : input_xyz, cons ? n_output : n_input, cons ? input_xyz : output_xyz, cons ? n_input : n_output, tags);
  });
}

int rbfmap_tag_mesh_second_round(rbfmap_handle *h, const double *mesh_xyz, int64_t n, const unsigned char *owner, unsigned char *tags)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    if (n < 0 || !tags || !owner)
      throw MapError(RBFMAP_ERR_INVALID, "invalid mesh / owner / tag buffer");
    h->impl->tagMeshSecondRound(mesh_xyz, n, owner, tags);
  });
}

int rbfmap_estimate_cluster_radius(rbfmap_handle *h, const double *mesh_xyz, int64_t n, const double bb_min[3], const double bb_max[3], double *radius)
{
  if (!h || !radius)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    if (n <= 0 || !mesh_xyz || !bb_min || !bb_max)
      throw MapError(RBFMAP_ERR_INVALID, "invalid mesh / bounding box");
    *radius = h->impl->estimateClusterRadiusHost(mesh_xyz, n, bb_min, bb_max);
  });
}

int rbfmap_nccl_unique_id(char id[128])
{
  return guarded(nullptr, [&] {
    ncclUniqueId u;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    RBF_NCCL(ncclApi().GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
  });
}

int rbfmap_distributed_init(rbfmap_handle *h, const char id[128], int rank, int world)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] { h->impl->distributedInit(id, rank, world); });
}

int rbfmap_partition_rows(int64_t n, int world, int rank, int64_t *begin, int64_t *end)
{
  if (n < 0 || world < 1 || rank < 0 || rank >= world || !begin || !end)
    return RBFMAP_ERR_INVALID;
  int64_t slice;
  partitionRows(n, world, rank, *begin, *end, slice);
  return RBFMAP_OK;
}

int rbfmap_partition_slabs(const int64_t *histogram, int n_bins, int world, int halo_bins, int *cuts)
{
  if (!histogram || !cuts || n_bins < 1 || world < 1 || halo_bins < 0)
    return RBFMAP_ERR_INVALID;
  std::vector<long long> h(histogram, histogram + n_bins);
  pumPartitionSlabs(h.data(), n_bins, world, halo_bins, cuts);
  return RBFMAP_OK;
}

int rbfmap_clear(rbfmap_handle *h)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(h, [&] {
    AllocStreamScope allocScope(h->impl->stream());
    h->impl->clear();
  });
}

int rbfmap_has_computed_mapping(const rbfmap_handle *h) { return h && h->impl->hasComputedMapping() ? 1 : 0; }

void rbfmap_destroy(rbfmap_handle *h)
{
  if (!h)
    return;
  DeviceScope scope(h->impl ? h->impl->device() : -1);
  delete h;
}

const char *rbfmap_get_name(const rbfmap_handle *h) { return h ? h->name.c_str() : ""; }

const char *rbfmap_last_error(const rbfmap_handle *h) { return h ? h->lastError.c_str() : g_createError.c_str(); }

int rbfmap_get_stats(const rbfmap_handle *h, rbfmap_stats *stats)
{
  if (!h || !stats)
    return RBFMAP_ERR_INVALID;
  *stats = h->impl->stats;
  return RBFMAP_OK;
}

int rbfmap_pum_get_clusters(const rbfmap_handle *h, double *centers, int64_t *in_offsets, int64_t *out_offsets)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(const_cast<rbfmap_handle *>(h), [&] {
    int64_t              nc  = 0;
    const double        *c   = nullptr;
    const PumMembership *mem = nullptr;
    if (!pumInspect(h->impl.get(), nc, c, mem))
      throw MapError(RBFMAP_ERR_INVALID, "not a partition-of-unity mapping");
    if (nc == 0)
      return;
    if (centers)
      RBF_CUDA(cudaMemcpy(centers, c, 3 * nc * sizeof(double), cudaMemcpyDeviceToHost));
    if (in_offsets)
      RBF_CUDA(cudaMemcpy(in_offsets, mem->inOff.p, (nc + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (out_offsets)
      RBF_CUDA(cudaMemcpy(out_offsets, mem->outOff.p, (nc + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  });
}

int rbfmap_pum_get_cluster_ids(const rbfmap_handle *h, int32_t *in_ids, int32_t *out_ids)
{
  if (!h)
    return RBFMAP_ERR_INVALID;
  return guarded(const_cast<rbfmap_handle *>(h), [&] {
    int64_t              nc  = 0;
    const double        *c   = nullptr;
    const PumMembership *mem = nullptr;
    if (!pumInspect(h->impl.get(), nc, c, mem))
      throw MapError(RBFMAP_ERR_INVALID, "not a partition-of-unity mapping");
    if (nc == 0)
      return;
    if (in_ids && mem->sumIn > 0)
      RBF_CUDA(cudaMemcpy(in_ids, mem->inIds.p, mem->sumIn * sizeof(int), cudaMemcpyDeviceToHost));
    if (out_ids && mem->sumOut > 0)
      RBF_CUDA(cudaMemcpy(out_ids, mem->outIds.p, mem->sumOut * sizeof(int), cudaMemcpyDeviceToHost));
  });
}

int rbfmap_eval_basis(int basis, double support_radius, double shape_parameter, const double *radii, int64_t n, double *values)
{
  return guarded(nullptr, [&] {
    const RbfParams f = makeRbfParams(basis, support_radius, shape_parameter);
    int             count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
      cudaGetLastError();
      throw MapError(RBFMAP_ERR_CUDA, "rbfmap: no CUDA device available; this backend has no CPU fallback.");
    }
    if (n <= 0)
      return;
    DevBuf<double> r, v;
    r.alloc(n);
    v.alloc(n);
    RBF_CUDA(cudaMemcpy(r.p, radii, n * sizeof(double), cudaMemcpyHostToDevice));
    evalBasisKernel<<<divUp(n, 256), 256>>>(f, r.p, n, v.p);
    RBF_CUDA(cudaGetLastError());
    RBF_CUDA(cudaMemcpy(values, v.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int rbfmap_debug_pair_basis(int basis, double support_radius, double shape_parameter, int dimensions, const double *a, const double *b, int64_t n, int batched,
                            double *values)
{
  return guarded(nullptr, [&] {
    const RbfParams f = makeRbfParams(basis, support_radius, shape_parameter);
    if (dimensions != 2 && dimensions != 3)
      throw MapError(RBFMAP_ERR_INVALID, "dimensions has to be 2 or 3");
    if (n <= 0)
      return;
    DevBuf<double> da, db, dv;
    da.alloc(3 * n);
    db.alloc(3 * n);
    dv.alloc(n);
    RBF_CUDA(cudaMemcpy(da.p, a, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    RBF_CUDA(cudaMemcpy(db.p, b, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    debugPairBasis(f, dimensions == 3, da.p, db.p, n, batched, dv.p);
    RBF_CUDA(cudaDeviceSynchronize());
    RBF_CUDA(cudaMemcpy(values, dv.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int rbfmap_debug_dist_sqrt(const double *x, int64_t n, double *out)
{
  return guarded(nullptr, [&] {
    if (n <= 0)
      return;
    DevBuf<double> a, b;
    a.alloc(n);
    b.alloc(n);
    RBF_CUDA(cudaMemcpy(a.p, x, n * sizeof(double), cudaMemcpyHostToDevice));
    debugDistSqrt(a.p, n, b.p);
    RBF_CUDA(cudaDeviceSynchronize());
    RBF_CUDA(cudaMemcpy(out, b.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int rbfmap_host_alloc(void **ptr, int64_t bytes)
{
  if (!ptr || bytes < 0)
    return RBFMAP_ERR_INVALID;
  *ptr = nullptr;
  return guarded(nullptr, [&] {
    if (bytes > 0)
      RBF_CUDA(cudaHostAlloc(ptr, static_cast<size_t>(bytes), cudaHostAllocPortable));
  });
}

void rbfmap_host_free(void *ptr)
{
  if (ptr)
    cudaFreeHost(ptr);
}

void rbfmap_struct_sizes(int *config_bytes, int *stats_bytes)
{
  if (config_bytes)
    *config_bytes = static_cast<int>(sizeof(rbfmap_config));
  if (stats_bytes)
    *stats_bytes = static_cast<int>(sizeof(rbfmap_stats));
}

int rbfmap_device_count(void)
{
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

int rbfmap_device_info(int device, int *sm_major, int *sm_minor, int *sm_count, int64_t *global_mem_bytes)
{
  cudaDeviceProp p{};
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) {
    cudaGetLastError();
    return RBFMAP_ERR_CUDA;
  }
  if (sm_major) *sm_major = p.major;
  if (sm_minor) *sm_minor = p.minor;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (global_mem_bytes) *global_mem_bytes = static_cast<int64_t>(p.totalGlobalMem);
  return RBFMAP_OK;
}

} // extern "C"
