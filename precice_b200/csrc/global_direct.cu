// global_direct.cu — rbf-global-direct on one B200: host driver + polynomial kernels.
//
// Mirrors mapping::RadialBasisFctMapping<RadialBasisFctSolver<RBF>> (mapping/RadialBasisFctMapping.hpp:94-389, serial
// branch; the multi-rank gather/scatter of :116-149, :208-307 stays host code of the caller) with
// RadialBasisFctSolver's constructor / solveConsistent / solveConservative (mapping/RadialBasisFctSolver.hpp:340-468).
// "The global-direct solve stays single-GPU" (BASELINE.json north_star): replicas only, no collective.
//
//  * strictly positive definite functions: C assembled tile-wise, blocked Cholesky with FP64 tensor-core trailing
//    updates (dense.cu), triangular solves per map call;
//  * the evaluation matrix A (m x n) is never stored: A lambda and A^T u are matrix-free products;
//  * separated polynomial: least-squares fit in the centred / scaled basis via its Gram matrix + one refinement step;
//    the cond > 1e5 axis-drop rule (RadialBasisFctSolver.hpp:383-402) is evaluated on the raw polynomial matrix.
#include "dense.cuh"
#include "mapping.cuh"
#include "small_dense.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace rbfmap {
namespace {

struct PolyGlobal {
  int    mode; // 0 none, 1 centred/scaled basis with inverse Gram matrix G
  int    mask; // active axes of the polynomial (bit d)
  int    p;
  double cx, cy, cz, scale;
  double G[16];
};

__device__ __forceinline__ void polyRowG(const PolyGlobal &pg, double x, double y, double z, double (&q)[4])
{
  q[0] = 1.0;
  q[1] = q[2] = q[3] = 0.0;
  const double v[3] = {(x - pg.cx) * pg.scale, (y - pg.cy) * pg.scale, (z - pg.cz) * pg.scale};
  int          k    = 1;
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if (pg.mask & (1 << d)) {
      if (k == 1) q[1] = v[d];
      else if (k == 2) q[2] = v[d];
      else q[3] = v[d];
      ++k;
    }
}

// raw moments sum q q^T, q = [1, x, y, z], and the axis extents (RadialBasisFctSolver.hpp:115-137, 383-402)
__global__ void momentsKernel(const double *__restrict__ xyz, long long n, double *__restrict__ m16, unsigned long long *__restrict__ ext6)
{
  double m[10];
#pragma unroll
  for (int k = 0; k < 10; ++k)
    m[k] = 0.0;
  double lo[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
  double hi[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double q[4] = {1.0, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    int          k    = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s)
        m[k++] += q[r] * q[s];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = fmin(lo[d], q[d + 1]);
      hi[d] = fmax(hi[d], q[d + 1]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 10; ++k)
      m[k] += __shfl_xor_sync(0xffffffffu, m[k], o);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
    int k = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s)
        atomicAdd(&m16[4 * r + s], m[k++]);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      atomicMin(&ext6[d], encodeOrdered(lo[d]));
      atomicMax(&ext6[3 + d], encodeOrdered(hi[d]));
    }
  }
}

__global__ void centredGramKernel(const double *__restrict__ xyz, long long n, PolyGlobal pg, double *__restrict__ g16)
{
  double m[10];
#pragma unroll
  for (int k = 0; k < 10; ++k)
    m[k] = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    double q[4];
    polyRowG(pg, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], q);
    int k = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s)
        m[k++] += q[r] * q[s];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 10; ++k)
      m[k] += __shfl_xor_sync(0xffffffffu, m[k], o);
  if ((threadIdx.x & 31) == 0) {
    int k = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s)
        atomicAdd(&g16[4 * r + s], m[k++]);
  }
}

// t[k * nrhs + c] += sum_i Q(i,k) (vals[c*ld + i] + sgn * (Q beta)(i, c))
__global__ void polyProjectKernel(const double *__restrict__ xyz, long long n, PolyGlobal pg, const double *__restrict__ vals, long long ld, int nrhs,
                                  const double *__restrict__ beta, double sgn, double *__restrict__ t, long long tk = -1, long long tc = 1)
{
  if (tk < 0)
    tk = nrhs; // default layout t[k * nrhs + c]
  for (int c = 0; c < nrhs; ++c) {
    double acc[4] = {0, 0, 0, 0};
    double b[4]   = {0, 0, 0, 0};
    if (beta)
      for (int k = 0; k < 4; ++k)
        b[k] = k < pg.p ? beta[k * nrhs + c] : 0.0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
      double q[4];
      polyRowG(pg, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], q);
      const double r = vals[c * ld + i] + sgn * (q[0] * b[0] + q[1] * b[1] + q[2] * b[2] + q[3] * b[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        acc[k] += q[k] * r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < pg.p)
          atomicAdd(&t[k * tk + c * tc], acc[k]);
  }
}

// beta += G (a - b)   (a, b: 4 x nrhs; b may be null)
__global__ void polySolveKernel(PolyGlobal pg, const double *__restrict__ a, const double *__restrict__ b, int nrhs, double *__restrict__ beta)
{
  const int c = threadIdx.x;
  if (c >= nrhs)
    return;
  double t[4];
  for (int k = 0; k < 4; ++k)
    t[k] = a[k * nrhs + c] - (b ? b[k * nrhs + c] : 0.0);
  for (int k = 0; k < 4; ++k) {
    double s = 0.0;
    for (int l = 0; l < 4; ++l)
      s += pg.G[4 * k + l] * t[l];
    beta[k * nrhs + c] += s;
  }
}

// vals[c*ld + i] += sgn * (Q beta)(i, c)
__global__ void polyApplyKernel(const double *__restrict__ xyz, long long n, PolyGlobal pg, const double *__restrict__ beta, int nrhs, double sgn,
                                double *__restrict__ vals, long long ld, long long bk = -1, long long bc = 1)
{
  if (bk < 0)
    bk = nrhs; // default layout beta[k * nrhs + c]
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  double q[4];
  polyRowG(pg, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], q);
  for (int c = 0; c < nrhs; ++c) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < pg.p)
        s += q[k] * beta[k * bk + c * bc];
    vals[c * ld + i] += sgn * s;
  }
}

// integrated polynomial (RadialBasisFctSolver.hpp:140-162): C(i, n+q) = C(n+q, i) = P(i, q), zero p x p block
__global__ void polyBlockKernel(const double *__restrict__ xyz, long long n, PolyGlobal pg, double *__restrict__ C, long long ld)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) {
    double q[4];
    polyRowG(pg, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], q);
    for (int k = 0; k < pg.p; ++k) {
      C[i + (n + k) * ld] = q[k];
      C[(n + k) + i * ld] = q[k];
    }
  }
  if (i < pg.p)
    for (int k = 0; k < pg.p; ++k)
      C[(n + i) + (n + k) * ld] = 0.0;
}

inline int reduceGrid(int64_t n) { return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(smCount()) * 8))); }

class GlobalDirectMapping final : public DeviceMapping {
public:
  explicit GlobalDirectMapping(const rbfmap_config &cfg)
      : DeviceMapping(cfg)
  {
    // RadialBasisFctMapping.hpp:90
    if (basisIsSpd(cfg.basis) && cfg.polynomial == RBFMAP_POLYNOMIAL_ON)
      throw MapError(RBFMAP_ERR_INVALID, "The integrated polynomial (polynomial=\"on\") is not supported for the selected radial-basis function. "
                                         "Please select another radial-basis function or change the polynomial configuration.");
    // RadialBasisFctBaseMapping.hpp:100-108
    bool anyAlive = false;
    for (int d = 0; d < cfg.dimensions; ++d)
      anyAlive |= (cfg.dead_axis[d] == 0);
    if (!anyAlive)
      throw MapError(RBFMAP_ERR_INVALID, "You cannot set all axes to dead for an RBF mapping. Please remove one of the respective mapping's "
                                         "\"x-dead\", \"y-dead\", or \"z-dead\" attributes.");
    _spd = basisIsSpd(cfg.basis);
    for (int d = 0; d < 3; ++d)
      _active[d] = d < cfg.dimensions && cfg.dead_axis[d] == 0;
  }

  std::string name() const override { return "global-direct RBF (cuda-executor)"; }

  void computeMapping(const double *dInputXyz, int64_t nInput, const double *dOutputXyz, int64_t nOutput) override;
  void mapConsistent(const double *dIn, int dims, double *dOut) override;
  void mapConservative(const double *dIn, int dims, double *dOut) override;
  void clear() override
  {
    _inXyz.release();
    _outXyz.release();
    _C.release();
    _invDiag.release();
    _piv.release();
    _computed = false;
    stats     = rbfmap_stats{};
  }

private:
  void     setupPolynomial();
  AxisMask mask() const { return AxisMask{_active[0] ? 1.0 : 0.0, _active[1] ? 1.0 : 0.0, _active[2] ? 1.0 : 0.0}; }

  void solveSystem(double *dB, int nrhs)
  {
    if (_spd)
      denseCholeskySolve(_C.p, _np, _invDiag.p, dB, _np, nrhs, _stream);
    else
      denseLuSolve(_C.p, _np, _piv.p, dB, _np, nrhs, _stream);
  }

  bool           _spd       = true;
  bool           _active[3] = {true, true, true};
  DevBuf<double> _inXyz, _outXyz;
  int64_t        _nIn = 0, _nOut = 0, _np = 0;
  DevBuf<double> _C, _invDiag;
  DevBuf<int>    _piv;
  PolyGlobal     _poly{};   // separated polynomial
  PolyGlobal     _polyOn{}; // integrated polynomial (mode 0 = none)
};

void GlobalDirectMapping::setupPolynomial()
{
  _poly        = PolyGlobal{};
  _poly.mode   = 0;
  _polyOn      = PolyGlobal{};
  _polyOn.mode = 0;
  if (_cfg.polynomial == RBFMAP_POLYNOMIAL_OFF)
    return;
  DevBuf<double>             m16;
  DevBuf<unsigned long long> ext;
  m16.alloc(16);
  ext.alloc(6);
  unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
  RBF_CUDA(cudaMemsetAsync(m16.p, 0, 16 * sizeof(double), _stream));
  RBF_CUDA(cudaMemcpyAsync(ext.p, init, sizeof(init), cudaMemcpyHostToDevice, _stream));
  momentsKernel<<<reduceGrid(_nIn), 256, 0, _stream>>>(_inXyz.p, _nIn, m16.p, ext.p);
  RBF_LAUNCHED();
  double             hm[16];
  unsigned long long he[6];
  RBF_CUDA(cudaMemcpyAsync(hm, m16.p, sizeof(hm), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaMemcpyAsync(he, ext.p, sizeof(he), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  double lo[3], hi[3];
  for (int d = 0; d < 3; ++d) {
    lo[d] = decodeOrdered(he[d]);
    hi[d] = decodeOrdered(he[3 + d]);
  }
  int mask = (_active[0] ? 1 : 0) | (_active[1] ? 2 : 0) | (_active[2] ? 4 : 0);
  if (_cfg.polynomial == RBFMAP_POLYNOMIAL_ON) {
    // integrated polynomial over all active axes (no axis dropping, RadialBasisFctSolver.hpp:140-162). The columns are
    // written in the centred / scaled basis: it spans the same space as [1, x_active...], so the interpolant is the
    // same function, while the augmented matrix is far better scaled.
    _polyOn.mode = 1;
    _polyOn.mask = mask;
    _polyOn.p    = 1 + ((mask & 1) ? 1 : 0) + ((mask & 2) ? 1 : 0) + ((mask & 4) ? 1 : 0);
    _polyOn.cx   = 0.5 * (lo[0] + hi[0]);
    _polyOn.cy   = 0.5 * (lo[1] + hi[1]);
    _polyOn.cz   = 0.5 * (lo[2] + hi[2]);
    double sp    = 0.0;
    for (int d = 0; d < 3; ++d)
      if (mask & (1 << d))
        sp = std::max(sp, hi[d] - lo[d]);
    _polyOn.scale = sp > 0.0 ? 2.0 / sp : 1.0;
    return;
  }
  // RadialBasisFctSolver.hpp:383-402: drop the active axis with the smallest extent while cond(Q) > 1e5
  int p    = 1;
  while (true) {
    p = 1 + ((mask & 1) ? 1 : 0) + ((mask & 2) ? 1 : 0) + ((mask & 4) ? 1 : 0);
    if (_nIn < p)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-global-direct: fewer vertices than polynomial parameters are not supported on the device.");
    double g[4][4];
    for (int r = 0; r < 4; ++r)
      for (int s = 0; s < 4; ++s) {
        const bool   ar = r == 0 || (mask & (1 << (r - 1))), as = s == 0 || (mask & (1 << (s - 1)));
        const double v  = r <= s ? hm[4 * r + s] : hm[4 * s + r];
        g[r][s]         = (ar && as) ? v : (r == s ? hm[0] : 0.0);
      }
    double evMin, evMax;
    jacobiEigenvalues4(g, evMin, evMax);
    const double cond = std::sqrt(std::max(evMax, 0.0)) / std::max(std::sqrt(std::max(evMin, 0.0)), kZeroDiff);
    if (cond > 1e5) {
      int    drop = -1;
      double best = 0.0;
      for (int d = 0; d < 3; ++d)
        if (mask & (1 << d)) {
          const double e = std::abs(hi[d] - lo[d]);
          if (drop < 0 || e < best) {
            best = e;
            drop = d;
          }
        }
      if (drop < 0)
        break;
      mask &= ~(1 << drop);
    } else {
      break;
    }
  }
  _poly.mode  = 1;
  _poly.mask  = mask;
  _poly.p     = p;
  _poly.cx    = 0.5 * (lo[0] + hi[0]);
  _poly.cy    = 0.5 * (lo[1] + hi[1]);
  _poly.cz    = 0.5 * (lo[2] + hi[2]);
  double span = 0.0;
  for (int d = 0; d < 3; ++d)
    if (mask & (1 << d))
      span = std::max(span, hi[d] - lo[d]);
  _poly.scale = span > 0.0 ? 2.0 / span : 1.0;
  RBF_CUDA(cudaMemsetAsync(m16.p, 0, 16 * sizeof(double), _stream));
  centredGramKernel<<<reduceGrid(_nIn), 256, 0, _stream>>>(_inXyz.p, _nIn, _poly, m16.p);
  RBF_LAUNCHED();
  RBF_CUDA(cudaMemcpyAsync(hm, m16.p, sizeof(hm), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  double g[4][4], inv[4][4];
  for (int r = 0; r < 4; ++r)
    for (int s = 0; s < 4; ++s)
      g[r][s] = r <= s ? hm[4 * r + s] : hm[4 * s + r];
  for (int r = p; r < 4; ++r)
    g[r][r] = 1.0;
  invertSpd4(g, inv);
  for (int r = 0; r < 4; ++r)
    for (int s = 0; s < 4; ++s)
      _poly.G[4 * r + s] = (r < p && s < p) ? inv[r][s] : 0.0;
}

void GlobalDirectMapping::computeMapping(const double *dInputXyz, int64_t nInput, const double *dOutputXyz, int64_t nOutput)
{
  g_launches = 0;
  AllocStreamScope allocScope(_stream);
  EventTimer       whole(_stream);
  whole.start();
  PhaseTrace trace("global.computeMapping", _stream);
  clear(); // the global variant simply replaces its solver (RadialBasisFctMapping.hpp:155-163)
  _nInput  = nInput;
  _nOutput = nOutput;
  // RadialBasisFctMapping.hpp:108-114
  const bool cons = isConservative();
  _nIn            = cons ? nOutput : nInput;
  _nOut           = cons ? nInput : nOutput;
  _inXyz.alloc(3 * std::max<int64_t>(_nIn, 1));
  _outXyz.alloc(3 * std::max<int64_t>(_nOut, 1));
  if (_nIn > 0)
    RBF_CUDA(cudaMemcpyAsync(_inXyz.p, cons ? dOutputXyz : dInputXyz, 3 * _nIn * sizeof(double), cudaMemcpyDeviceToDevice, _stream));
  if (_nOut > 0)
    RBF_CUDA(cudaMemcpyAsync(_outXyz.p, cons ? dInputXyz : dOutputXyz, 3 * _nOut * sizeof(double), cudaMemcpyDeviceToDevice, _stream));
  stats.n_in  = _nIn;
  stats.n_out = _nOut;
  if (_nIn == 0) {
    _computed = true;
    return;
  }
  if (_nIn > 200000)
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-global-direct: the dense system does not fit one device (use rbf-pum-direct or rbf-global-iterative).");
  setupPolynomial();
  trace.phase("polynomial");
  const int64_t nSys = _nIn + (_polyOn.mode ? _polyOn.p : 0);
  if (_polyOn.mode && _nIn < _polyOn.p) // PRECICE_ASSERT(inputSize >= 1 + polyparams), RadialBasisFctSolver.hpp:180
    throw MapError(RBFMAP_ERR_INVALID, "The integrated polynomial needs at least as many vertices as polynomial parameters.");
  _np = padTo(nSys, kPanelNb);
  _C.alloc(static_cast<size_t>(_np) * _np + 256);
  trace.phase("copy + alloc");
  EventTimer tf(_stream);
  tf.start();
  bool ok;
  if (_spd) {
    _invDiag.alloc(static_cast<size_t>(_np / kDiagNb) * kDiagNb * kDiagNb);
    denseAssemble(_rbf, mask(), _inXyz.p, _nIn, _C.p, _np, true, _stream);
    trace.phase("assemble C");
    ok = denseCholesky(_C.p, _np, _invDiag.p, _stream);
  } else {
    _piv.alloc(_np);
    denseAssemble(_rbf, mask(), _inXyz.p, _nIn, _C.p, _np, false, _stream);
    if (_polyOn.mode) {
      polyBlockKernel<<<divUp(_nIn, 256), 256, 0, _stream>>>(_inXyz.p, _nIn, _polyOn, _C.p, _np);
      RBF_LAUNCHED();
    }
    trace.phase("assemble C");
    ok = denseLu(_C.p, _np, _piv.p, _stream);
  }
  stats.ms_assembly_factor = tf.stop();
  trace.phase("factorisation");
  if (!ok) {
    // RadialBasisFctSolver.hpp:360-365
    const char *inName = cons ? "output" : "input", *outName = cons ? "input" : "output";
    throw MapError(RBFMAP_ERR_SINGULAR, std::string("The interpolation matrix of the RBF mapping from mesh \"") + inName + "\" to mesh \"" + outName +
                                            "\" is not invertable. This means that the mapping problem is not well-posed. "
                                            "Please check if your coupling meshes are correct (e.g. no vertices are duplicated) or reconfigure "
                                            "your basis-function (e.g. reduce the support-radius).");
  }
  stats.device_bytes       = static_cast<int64_t>(_C.bytes() + _invDiag.bytes() + _inXyz.bytes() + _outXyz.bytes());
  stats.ms_compute_kernels = whole.stop();
  stats.kernel_launches    = g_launches;
  _computed                = true;
}

// solveConsistent: RadialBasisFctSolver.hpp:441-468 ; caller side RadialBasisFctMapping.hpp:310-389
void GlobalDirectMapping::mapConsistent(const double *dIn, int dims, double *dOut)
{
  if (!_computed)
    throw MapError(RBFMAP_ERR_STATE, "The mapping has not been computed; call computeMapping() first.");
  const int        launchesBefore = g_launches;
  AllocStreamScope allocScope(_stream);
  EventTimer       t(_stream);
  t.start();
  if (_nOut > 0)
    RBF_CUDA(cudaMemsetAsync(dOut, 0, static_cast<size_t>(_nOut) * dims * sizeof(double), _stream));
  if (_nIn > 0 && _nOut > 0) {
    DevBuf<double> f, o, beta, tt;
    f.alloc(static_cast<size_t>(_np) * dims);
    o.alloc(static_cast<size_t>(_nOut) * dims);
    interleavedToPlanar(dIn, _nIn, dims, f.p, _np, _np, _stream);
    if (_poly.mode == 1) {
      beta.alloc(4 * dims);
      tt.alloc(4 * dims);
      RBF_CUDA(cudaMemsetAsync(beta.p, 0, 4 * dims * sizeof(double), _stream));
      for (int pass = 0; pass < 2; ++pass) { // normal equations + one refinement step
        RBF_CUDA(cudaMemsetAsync(tt.p, 0, 4 * dims * sizeof(double), _stream));
        polyProjectKernel<<<reduceGrid(_nIn), 256, 0, _stream>>>(_inXyz.p, _nIn, _poly, f.p, _np, dims, beta.p, -1.0, tt.p);
        polySolveKernel<<<1, std::max(32, dims), 0, _stream>>>(_poly, tt.p, nullptr, dims, beta.p);
        RBF_LAUNCHED();
        RBF_LAUNCHED();
      }
      polyApplyKernel<<<divUp(_nIn, 256), 256, 0, _stream>>>(_inXyz.p, _nIn, _poly, beta.p, dims, -1.0, f.p, _np);
      RBF_LAUNCHED();
    }
    solveSystem(f.p, dims);
    denseEvaluate(_rbf, mask(), _outXyz.p, _nOut, _inXyz.p, _nIn, f.p, _np, o.p, _nOut, dims, false, _stream);
    if (_polyOn.mode) { // out += P_out lambda_poly (the last p rows of the solution, RadialBasisFctSolver.hpp:238-240)
      polyApplyKernel<<<divUp(_nOut, 256), 256, 0, _stream>>>(_outXyz.p, _nOut, _polyOn, f.p + _nIn, dims, 1.0, o.p, _nOut, 1, _np);
      RBF_LAUNCHED();
    }
    if (_poly.mode == 1) {
      polyApplyKernel<<<divUp(_nOut, 256), 256, 0, _stream>>>(_outXyz.p, _nOut, _poly, beta.p, dims, 1.0, o.p, _nOut);
      RBF_LAUNCHED();
    }
    planarToInterleaved(o.p, _nOut, _nOut, dims, dOut, _stream);
    RBF_CUDA(cudaGetLastError());
  }
  stats.ms_map_kernels = t.stop();
  stats.kernel_launches += g_launches - launchesBefore;
}

// solveConservative: RadialBasisFctSolver.hpp:413-438 ; caller side RadialBasisFctMapping.hpp:199-308
void GlobalDirectMapping::mapConservative(const double *dIn, int dims, double *dOut)
{
  if (!_computed)
    throw MapError(RBFMAP_ERR_STATE, "The mapping has not been computed; call computeMapping() first.");
  const int        launchesBefore = g_launches;
  AllocStreamScope allocScope(_stream);
  EventTimer       t(_stream);
  t.start();
  // values live on the solver-side outMesh (= Mapping::input()), the result on the solver-side inMesh
  if (_nIn > 0)
    RBF_CUDA(cudaMemsetAsync(dOut, 0, static_cast<size_t>(_nIn) * dims * sizeof(double), _stream));
  if (_nIn > 0 && _nOut > 0) {
    DevBuf<double> u, mu, eps, tt, gamma;
    u.alloc(static_cast<size_t>(_nOut) * dims);
    mu.alloc(static_cast<size_t>(_np) * dims);
    interleavedToPlanar(dIn, _nOut, dims, u.p, _nOut, _nOut, _stream);
    RBF_CUDA(cudaMemsetAsync(mu.p, 0, static_cast<size_t>(_np) * dims * sizeof(double), _stream));
    // Au = A^T u
    denseEvaluate(_rbf, mask(), _inXyz.p, _nIn, _outXyz.p, _nOut, u.p, _nOut, mu.p, _np, dims, false, _stream);
    if (_polyOn.mode) { // the polynomial rows of A^T u: P_out^T u
      polyProjectKernel<<<reduceGrid(_nOut), 256, 0, _stream>>>(_outXyz.p, _nOut, _polyOn, u.p, _nOut, dims, nullptr, 0.0, mu.p + _nIn, 1, _np);
      RBF_LAUNCHED();
    }
    solveSystem(mu.p, dims);
    if (_poly.mode == 1) {
      eps.alloc(4 * dims);
      tt.alloc(4 * dims);
      gamma.alloc(4 * dims);
      RBF_CUDA(cudaMemsetAsync(eps.p, 0, 4 * dims * sizeof(double), _stream));
      RBF_CUDA(cudaMemsetAsync(gamma.p, 0, 4 * dims * sizeof(double), _stream));
      polyProjectKernel<<<reduceGrid(_nOut), 256, 0, _stream>>>(_outXyz.p, _nOut, _poly, u.p, _nOut, dims, nullptr, 0.0, eps.p); // V^T u
      RBF_LAUNCHED();
      for (int pass = 0; pass < 2; ++pass) { // gamma += G (eps - Q^T (mu + Q gamma))
        RBF_CUDA(cudaMemsetAsync(tt.p, 0, 4 * dims * sizeof(double), _stream));
        polyProjectKernel<<<reduceGrid(_nIn), 256, 0, _stream>>>(_inXyz.p, _nIn, _poly, mu.p, _np, dims, gamma.p, 1.0, tt.p);
        polySolveKernel<<<1, std::max(32, dims), 0, _stream>>>(_poly, eps.p, tt.p, dims, gamma.p);
        RBF_LAUNCHED();
        RBF_LAUNCHED();
      }
      polyApplyKernel<<<divUp(_nIn, 256), 256, 0, _stream>>>(_inXyz.p, _nIn, _poly, gamma.p, dims, 1.0, mu.p, _np);
      RBF_LAUNCHED();
    }
    planarToInterleaved(mu.p, _np, _nIn, dims, dOut, _stream);
    RBF_CUDA(cudaGetLastError());
  }
  stats.ms_map_kernels = t.stop();
  stats.kernel_launches += g_launches - launchesBefore;
}

} // namespace

std::unique_ptr<DeviceMapping> makeGlobalDirectMapping(const rbfmap_config &cfg) { return std::make_unique<GlobalDirectMapping>(cfg); }

} // namespace rbfmap
