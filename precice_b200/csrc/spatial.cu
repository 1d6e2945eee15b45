// spatial.cu — device implementations of the neighbour-search primitives declared in spatial.cuh.
#include "spatial.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace rbfmap {
namespace {

constexpr int kBlock = 256;

inline int gridFor(int64_t n, int perThread = 1)
{
  const int64_t want = (n + static_cast<int64_t>(kBlock) * perThread - 1) / (static_cast<int64_t>(kBlock) * perThread);
  const int64_t cap  = static_cast<int64_t>(smCount()) * 16; // grid-stride loops; a multiple of the SM count
  return static_cast<int>(std::max<int64_t>(1, std::min(want, cap)));
}

// ------------------------------------------------------------------ device-to-device copy
// cudaMemcpyAsync(device -> device) of a 48 MB mesh took ~0.4 ms on this B200 when the GPU had been idle for a moment
// (copy engine), against ~0.03 ms for a kernel.
__global__ void copyKernel(const double2 *__restrict__ src, double2 *__restrict__ dst, int64_t n2, const double *__restrict__ tailSrc,
                           double *__restrict__ tailDst, int tail)
{
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n2; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
  if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < tail)
    tailDst[threadIdx.x] = tailSrc[threadIdx.x];
}
__global__ void copyScalarKernel(const double *__restrict__ src, double *__restrict__ dst, int64_t n)
{
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

// ------------------------------------------------------------------ bounds
// The coordinates are read as one flat array of 3 n doubles with a stride that is a multiple of 3 (192 threads per block):
// every thread stays on one axis and a warp reads 256 consecutive bytes per load, instead of three loads 24 bytes apart
// per vertex. Min / max are combined per block in shared memory (ordered-integer atomics) before the six global atomics.
constexpr int kBoundsBlock = 192;
__global__ void __launch_bounds__(kBoundsBlock) boundsKernel(const double *__restrict__ xyz, int64_t n, unsigned long long *__restrict__ keys /*6: min xyz, max xyz*/)
{
  __shared__ unsigned long long blockKeys[6];
  if (threadIdx.x < 6)
    blockKeys[threadIdx.x] = threadIdx.x < 3 ? ~0ull : 0ull;
  __syncthreads();
  const int64_t total  = 3 * n;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kBoundsBlock; // multiple of 3
  double        lo = DBL_MAX, hi = -DBL_MAX;
  int64_t       e = blockIdx.x * static_cast<int64_t>(kBoundsBlock) + threadIdx.x;
  for (; e + 3 * stride < total; e += 4 * stride) { // four independent loads in flight
    const double v0 = xyz[e], v1 = xyz[e + stride], v2 = xyz[e + 2 * stride], v3 = xyz[e + 3 * stride];
    lo = fmin(fmin(lo, v0), fmin(fmin(v1, v2), v3));
    hi = fmax(fmax(hi, v0), fmax(fmax(v1, v2), v3));
  }
  for (; e < total; e += stride) {
    const double v = xyz[e];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  const int d = threadIdx.x % 3; // kBoundsBlock and the stride are multiples of 3: the axis of every element this thread saw
  atomicMin(&blockKeys[d], encodeOrdered(lo)); // (DBL_MAX, -DBL_MAX) for a thread without elements, like an empty mesh
  atomicMax(&blockKeys[3 + d], encodeOrdered(hi));
  __syncthreads();
  if (threadIdx.x < 3)
    atomicMin(&keys[threadIdx.x], blockKeys[threadIdx.x]);
  else if (threadIdx.x < 6)
    atomicMax(&keys[threadIdx.x], blockKeys[threadIdx.x]);
}

// ------------------------------------------------------------------ nearest vertex to a few probe points
constexpr int kMaxProbes = 8;
struct Probes {
  double x[kMaxProbes], y[kMaxProbes], z[kMaxProbes];
  double r2[kMaxProbes]; // per-probe squared search radius (collectWithinKernel); negative = skip
  int    n;
};

__global__ void probeMinDistKernel(const double *__restrict__ xyz, int64_t n, Probes pr, unsigned long long *__restrict__ best)
{
  double b[kMaxProbes];
#pragma unroll
  for (int p = 0; p < kMaxProbes; ++p)
    b[p] = DBL_MAX;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
#pragma unroll
    for (int p = 0; p < kMaxProbes; ++p)
      if (p < pr.n)
        b[p] = fmin(b[p], sqDist3(pr.x[p], pr.y[p], pr.z[p], x, y, z));
  }
#pragma unroll
  for (int p = 0; p < kMaxProbes; ++p) {
    if (p < pr.n) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        b[p] = fmin(b[p], __shfl_xor_sync(0xffffffffu, b[p], o));
      if ((threadIdx.x & 31) == 0)
        atomicMin(&best[p], encodeOrdered(b[p]));
    }
  }
}

__global__ void probeArgMinKernel(const double *__restrict__ xyz, int64_t n, Probes pr, const unsigned long long *__restrict__ best, int *__restrict__ ids)
{
  double b[kMaxProbes];
#pragma unroll
  for (int p = 0; p < kMaxProbes; ++p)
    b[p] = p < pr.n ? decodeOrdered(best[p]) : 0.0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
#pragma unroll
    for (int p = 0; p < kMaxProbes; ++p)
      if (p < pr.n && sqDist3(pr.x[p], pr.y[p], pr.z[p], x, y, z) == b[p])
        atomicMin(&ids[p], static_cast<int>(i));
  }
}

// ------------------------------------------------------------------ k-th neighbour distance
__global__ void collectWithinKernel(const double *__restrict__ xyz, int64_t n, Probes pr, double *__restrict__ cand, int cap, int *__restrict__ counts)
{
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
#pragma unroll
    for (int p = 0; p < kMaxProbes; ++p)
      if (p < pr.n) {
        // RadialBasisFctSolver.hpp:102-113 argument order: (mesh vertex, sample vertex)
        const double d2 = sqDist3(x, y, z, pr.x[p], pr.y[p], pr.z[p]);
        if (d2 <= pr.r2[p]) {
          const int slot = atomicAdd(&counts[p], 1);
          if (slot < cap)
            cand[static_cast<size_t>(p) * cap + slot] = d2;
        }
      }
  }
}

// ------------------------------------------------------------------ binning
__global__ void binCountKernel(const double *__restrict__ xyz, int64_t n, BinGridView g, int *__restrict__ cellOf, int *__restrict__ counts, int regionAxis,
                               double regionLo, double regionHi)
{
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (regionAxis >= 0) { // sharded mapping: points outside this rank's region are not binned
      const double t = xyz[3 * i + regionAxis];
      if (!(t >= regionLo && t <= regionHi)) {
        cellOf[i] = -1;
        continue;
      }
    }
    int cx = cellCoordUnclamped(xyz[3 * i], g.ox, g.ihx, g.nx);
    int cy = cellCoordUnclamped(xyz[3 * i + 1], g.oy, g.ihy, g.ny);
    int cz = cellCoordUnclamped(xyz[3 * i + 2], g.oz, g.ihz, g.nz);
    cx     = min(max(cx, 0), g.nx - 1);
    cy     = min(max(cy, 0), g.ny - 1);
    cz     = min(max(cz, 0), g.nz - 1);
    const int c = (cx * g.ny + cy) * g.nz + cz;
    cellOf[i]   = c;
    atomicAdd(&counts[c], 1);
  }
}

__global__ void binScatterKernel(int64_t n, const int *__restrict__ cellOf, const int *__restrict__ cellStart, int *__restrict__ cursor, int *__restrict__ ids)
{
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c               = cellOf[i];
    if (c < 0)
      continue;
    const int slot            = atomicAdd(&cursor[c], 1);
    ids[cellStart[c] + slot]  = static_cast<int>(i);
  }
}

// ascending vertex IDs inside every cell -> the binned order is deterministic
__global__ void binSortCellsKernel(int64_t nCells, const int *__restrict__ cellStart, int *__restrict__ ids)
{
  for (int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; c < nCells; c += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int beg = cellStart[c], end = cellStart[c + 1];
    for (int a = beg + 1; a < end; ++a) {
      const int v = ids[a];
      int       b = a - 1;
      while (b >= beg && ids[b] > v) {
        ids[b + 1] = ids[b];
        --b;
      }
      ids[b + 1] = v;
    }
  }
}

__global__ void binGatherKernel(const double *__restrict__ xyz, int64_t n, const int *__restrict__ ids, double *__restrict__ binned)
{
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t v   = ids[i];
    binned[3 * i]     = xyz[3 * v];
    binned[3 * i + 1] = xyz[3 * v + 1];
    binned[3 * i + 2] = xyz[3 * v + 2];
  }
}

Probes makeProbes(const double *pts, int nProbes)
{
  if (nProbes > kMaxProbes)
    throw MapError(RBFMAP_ERR_INVALID, "too many probe points");
  Probes pr{};
  pr.n = nProbes;
  for (int p = 0; p < nProbes; ++p) {
    pr.x[p] = pts[3 * p];
    pr.y[p] = pts[3 * p + 1];
    pr.z[p] = pts[3 * p + 2];
  }
  return pr;
}

} // namespace

void deviceCopy(double *dst, const double *src, int64_t n, cudaStream_t s)
{
  if (n <= 0)
    return;
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    const int64_t n2 = n / 2;
    copyKernel<<<gridFor(std::max<int64_t>(n2, 1), 2), kBlock, 0, s>>>(reinterpret_cast<const double2 *>(src), reinterpret_cast<double2 *>(dst), n2, src + 2 * n2,
                                                                       dst + 2 * n2, static_cast<int>(n - 2 * n2));
  } else {
    copyScalarKernel<<<gridFor(n, 4), kBlock, 0, s>>>(src, dst, n);
  }
  RBF_LAUNCHED();
}

void computeBounds(const double *dXyz, int64_t n, double bbMin[3], double bbMax[3], cudaStream_t s)
{
  DevBuf<unsigned long long> keys;
  keys.alloc(6);
  unsigned long long init[6];
  for (int d = 0; d < 3; ++d) {
    init[d]     = ~0ull;
    init[3 + d] = 0ull;
  }
  RBF_CUDA(cudaMemcpyAsync(keys.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
  const int64_t wantBlocks = (3 * n + 8 * kBoundsBlock - 1) / (8 * kBoundsBlock);
  const int     grid       = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(wantBlocks, static_cast<int64_t>(smCount()) * 10)));
  boundsKernel<<<grid, kBoundsBlock, 0, s>>>(dXyz, n, keys.p);
  RBF_LAUNCHED();
  unsigned long long out[6];
  RBF_CUDA(cudaMemcpyAsync(out, keys.p, sizeof(out), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  for (int d = 0; d < 3; ++d) {
    bbMin[d] = decodeOrdered(out[d]);
    bbMax[d] = decodeOrdered(out[3 + d]);
  }
}

namespace {
__global__ void probeCoordsKernel(const double *__restrict__ xyz, const int *__restrict__ ids, int nProbes, double *__restrict__ out)
{
  const int t = threadIdx.x;
  if (t < 3 * nProbes)
    out[t] = xyz[3 * static_cast<long long>(ids[t / 3]) + t % 3];
}
} // namespace

void nearestVertices(const double *dXyz, int64_t n, const double *probes, int nProbes, int *nearestIds, cudaStream_t s, double *nearestXyz)
{
  const Probes               pr = makeProbes(probes, nProbes);
  DevBuf<unsigned long long> best;
  DevBuf<int>                ids;
  best.alloc(kMaxProbes);
  ids.alloc(kMaxProbes);
  RBF_CUDA(cudaMemsetAsync(best.p, 0xff, kMaxProbes * sizeof(unsigned long long), s));
  RBF_CUDA(cudaMemsetAsync(ids.p, 0x7f, kMaxProbes * sizeof(int), s));
  probeMinDistKernel<<<gridFor(n, 4), kBlock, 0, s>>>(dXyz, n, pr, best.p);
  RBF_LAUNCHED();
  probeArgMinKernel<<<gridFor(n, 4), kBlock, 0, s>>>(dXyz, n, pr, best.p, ids.p);
  RBF_LAUNCHED();
  RBF_CUDA(cudaMemcpyAsync(nearestIds, ids.p, nProbes * sizeof(int), cudaMemcpyDeviceToHost, s));
  DevBuf<double> coords;
  if (nearestXyz && n > 0) { // the coordinates of the found vertices travel with their IDs (one synchronisation)
    coords.alloc(3 * kMaxProbes);
    probeCoordsKernel<<<1, 32, 0, s>>>(dXyz, ids.p, nProbes, coords.p);
    RBF_LAUNCHED();
    RBF_CUDA(cudaMemcpyAsync(nearestXyz, coords.p, 3 * nProbes * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  RBF_CUDA(cudaStreamSynchronize(s));
}

void kthNeighbourDistance(const double *dXyz, int64_t n, const int *probeIds, int nProbes, int k, double guess, double *result, cudaStream_t s,
                          const double *probeXyz)
{
  k = static_cast<int>(std::min<int64_t>(k, n));
  if (k < 1)
    throw MapError(RBFMAP_ERR_INVALID, "vertices-per-cluster has to be greater zero.");
  // coordinates of the probe vertices
  std::vector<double> pts(3 * static_cast<size_t>(nProbes));
  if (probeXyz) {
    std::copy(probeXyz, probeXyz + 3 * nProbes, pts.begin());
  } else {
    for (int p = 0; p < nProbes; ++p)
      RBF_CUDA(cudaMemcpyAsync(&pts[3 * p], dXyz + 3 * static_cast<int64_t>(probeIds[p]), 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    RBF_CUDA(cudaStreamSynchronize(s));
  }
  Probes pr = makeProbes(pts.data(), nProbes);

  const int      cap = std::max(4096, 8 * k);
  DevBuf<double> cand;
  DevBuf<int>    counts;
  cand.alloc(static_cast<size_t>(kMaxProbes) * cap);
  counts.alloc(kMaxProbes);
  // the head of every probe's candidate list comes back with the counts (one synchronisation per round); the starting
  // radius holds ~2 k vertices, longer lists are fetched separately
  const int           head = std::min(cap, std::max(512, 4 * k));
  std::vector<double> host(static_cast<size_t>(kMaxProbes) * head), longList;
  std::vector<char>   done(nProbes, 0);
  // every probe searches with its own radius: grow while fewer than k vertices are inside, shrink (bisect) when the
  // candidate buffer overflows
  std::vector<double> radius(nProbes, guess > 0 && std::isfinite(guess) ? guess : 1.0), lo(nProbes, 0.0),
      hi(nProbes, std::numeric_limits<double>::infinity());
  int nDone = 0;
  for (int iter = 0; iter < 400 && nDone < nProbes; ++iter) {
    for (int p = 0; p < nProbes; ++p)
      pr.r2[p] = done[p] ? -1.0 : radius[p] * radius[p];
    RBF_CUDA(cudaMemsetAsync(counts.p, 0, kMaxProbes * sizeof(int), s));
    collectWithinKernel<<<gridFor(n, 4), kBlock, 0, s>>>(dXyz, n, pr, cand.p, cap, counts.p);
    RBF_LAUNCHED();
    int cnt[kMaxProbes];
    RBF_CUDA(cudaMemcpyAsync(cnt, counts.p, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    RBF_CUDA(cudaMemcpy2DAsync(host.data(), head * sizeof(double), cand.p, cap * sizeof(double), head * sizeof(double), nProbes, cudaMemcpyDeviceToHost, s));
    RBF_CUDA(cudaStreamSynchronize(s));
    for (int p = 0; p < nProbes; ++p) {
      if (done[p])
        continue;
      if (cnt[p] > cap) {
        hi[p]     = radius[p];
        radius[p] = lo[p] > 0 ? 0.5 * (lo[p] + hi[p]) : 0.5 * radius[p];
      } else if (cnt[p] < k) {
        lo[p]     = radius[p];
        radius[p] = std::isfinite(hi[p]) ? 0.5 * (lo[p] + hi[p]) : 2.0 * radius[p];
      } else {
        double *list = host.data() + static_cast<size_t>(p) * head;
        if (cnt[p] > head) {
          longList.resize(cnt[p]);
          RBF_CUDA(cudaMemcpy(longList.data(), cand.p + static_cast<size_t>(p) * cap, cnt[p] * sizeof(double), cudaMemcpyDeviceToHost));
          list = longList.data();
        }
        std::nth_element(list, list + (k - 1), list + cnt[p]);
        result[p] = std::sqrt(list[k - 1]);
        done[p]   = 1;
        ++nDone;
      }
    }
  }
  if (nDone < nProbes)
    throw MapError(RBFMAP_ERR_INVALID, "cluster radius estimation did not converge");
}

void BinGrid::build(const double *dXyz, int64_t nPts, const double bbMin[3], const double bbMax[3], double minCell, cudaStream_t s, int regionAxis,
                    double regionLo, double regionHi)
{
  n = nPts;
  // cell edge slightly above the query radius so that the 3x3x3 neighbourhood is conservative under rounding
  double cell = minCell * (1.0 + 1e-9);
  if (!(cell > 0) || !std::isfinite(cell))
    cell = 1.0;
  const double  capCells = static_cast<double>(std::min<int64_t>(std::max<int64_t>(8 * nPts, 1 << 20), int64_t(1) << 27));
  for (int it = 0; it < 200; ++it) {
    double total = 1.0;
    for (int d = 0; d < 3; ++d) {
      const double e  = std::max(0.0, bbMax[d] - bbMin[d]);
      const double c  = std::floor(e / cell) + 1.0;
      dims[d]         = static_cast<int>(std::min(c, 2097152.0));
      total *= static_cast<double>(dims[d]);
    }
    if (total <= capCells)
      break;
    cell *= 1.26;
  }
  for (int d = 0; d < 3; ++d) {
    origin[d] = bbMin[d];
    h[d]      = cell;
  }
  const int64_t nCells = static_cast<int64_t>(dims[0]) * dims[1] * dims[2];
  cellStart.alloc(nCells + 1);
  ids.alloc(std::max<int64_t>(nPts, 1));
  xyz.alloc(3 * std::max<int64_t>(nPts, 1));
  DevBuf<int> cellOf, counts;
  cellOf.alloc(std::max<int64_t>(nPts, 1));
  counts.alloc(nCells);
  RBF_CUDA(cudaMemsetAsync(counts.p, 0, nCells * sizeof(int), s));
  const BinGridView g = view();
  if (nPts > 0) {
    binCountKernel<<<gridFor(nPts), kBlock, 0, s>>>(dXyz, nPts, g, cellOf.p, counts.p, regionAxis, regionLo, regionHi);
    RBF_LAUNCHED();
  }
  exclusiveScan(counts.p, cellStart.p, nCells, s);
  if (regionAxis >= 0 && nPts > 0) { // number of points inside the region
    int binned = 0;
    RBF_CUDA(cudaMemcpyAsync(&binned, cellStart.p + nCells, sizeof(int), cudaMemcpyDeviceToHost, s));
    RBF_CUDA(cudaStreamSynchronize(s));
    n = binned;
  }
  if (nPts > 0) {
    RBF_CUDA(cudaMemsetAsync(counts.p, 0, nCells * sizeof(int), s));
    binScatterKernel<<<gridFor(nPts), kBlock, 0, s>>>(nPts, cellOf.p, cellStart.p, counts.p, ids.p);
    RBF_LAUNCHED();
    binSortCellsKernel<<<gridFor(nCells), kBlock, 0, s>>>(nCells, cellStart.p, ids.p);
    RBF_LAUNCHED();
    if (n > 0) {
      binGatherKernel<<<gridFor(n), kBlock, 0, s>>>(dXyz, n, ids.p, xyz.p);
      RBF_LAUNCHED();
    }
  }
  RBF_CUDA(cudaGetLastError()); // temporaries are released stream-ordered
}

} // namespace rbfmap
