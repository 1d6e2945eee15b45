// pum_cluster.cu — device side of the partition-of-unity clustering and vertex/cluster association.
//
// Replaces (on the device) the host R-tree driven parts of
//   impl/CreateClustering.hpp:132-140  tagEmptyClusters
//   impl/CreateClustering.hpp:150-160  projectClusterCentersToinputMesh
//   impl/CreateClustering.hpp:96-122, 173-195  getNeighborsWithinSphere / tagDuplicateCenters
//   impl/CreateClustering.hpp:205-208  removeTaggedVertices
//   impl/SphericalVertexCluster.hpp:153-162  per-cluster sphere queries + sorted ID sets
//   PartitionOfUnityMapping.hpp:290-341  computeNormalizedWeight (the per-vertex weight sums, accumulated from the cluster side)
#include "pum_device.cuh"

#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>

namespace rbfmap {
namespace {

// The sphere queries of the reference are strict comparisons of the ROUNDED distance: sqrt(d2) < r (query/Index.cpp:235).
// IEEE square root is monotone, so {d2 : sqrt(d2) < r} = {d2 <= t} with t the largest double whose root is below r; the
// kernels compare squared distances with t instead of taking a correctly rounded square root per candidate (a dozen
// FP64 instructions in kernels that are issue-bound).
inline double strictRadiusBound(double r)
{
  if (!(r > 0.0))
    return -1.0; // no squared distance qualifies
  double t = r * r;
  while (std::sqrt(t) >= r)
    t = std::nextafter(t, 0.0);
  while (std::sqrt(std::nextafter(t, std::numeric_limits<double>::infinity())) < r)
    t = std::nextafter(t, std::numeric_limits<double>::infinity());
  return t;
}

constexpr int kWarpsPerBlock = 8;
constexpr int kBlockThreads  = kWarpsPerBlock * 32;

inline int warpGrid(int64_t nWarps) { return static_cast<int>(std::max<int64_t>(1, (nWarps + kWarpsPerBlock - 1) / kWarpsPerBlock)); }

// Block-local compaction of the live (untagged) centres. On a surface mesh the tentative lattice is almost empty after
// the first test (2M-vertex surface: 12.7M lattice centres, 2-3 % alive); one warp per lattice centre would spend its time
// launching warps that read one tag and leave. Every block looks at kBlockThreads consecutive centres (one coalesced tag
// load), packs the live ones into shared memory and its warps then walk that list: body(c, lane), all 32 lanes together.
template <typename F>
__device__ __forceinline__ void forLiveCentres(const int *__restrict__ tag, int64_t total, F &&body)
{
  __shared__ int live[kBlockThreads];
  __shared__ int nLive;
  const int64_t  base = blockIdx.x * static_cast<int64_t>(kBlockThreads);
  const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    nLive = 0;
  __syncthreads();
  const int64_t  c     = base + threadIdx.x;
  const bool     alive = c < total && tag[c] == 0;
  const unsigned m     = __ballot_sync(0xffffffffu, alive);
  int            slot  = 0;
  if (lane == 0 && m != 0)
    slot = atomicAdd(&nLive, __popc(m));
  slot = __shfl_sync(0xffffffffu, slot, 0);
  if (alive)
    live[slot + __popc(m & ((1u << lane) - 1u))] = threadIdx.x;
  __syncthreads();
  const int n = nLive;
  for (int i = warp; i < n; i += kWarpsPerBlock)
    body(base + live[i], lane);
}

// ---- tagEmptyClusters on arbitrary (projected) centres: one warp per live centre -------------
// is any binned vertex strictly inside the sphere (squared-distance bound r2Bound, see strictRadiusBound)? The 3x3x3 cells
// around the centre are searched first: they hold the vertices closest to the centre, and in a populated region the
// first batch of 32 candidates already contains a hit; only otherwise the full 5x5x5 neighbourhood is walked.
__device__ __forceinline__ bool anyVertexInside(const BinGridView &g, double cx, double cy, double cz, double r2Bound, int lane)
{
  bool found = false;
  auto test  = [&](int pos, bool valid) {
    bool hit = false;
    if (valid)
      hit = sqDist3(cx, cy, cz, g.xyz[3 * pos], g.xyz[3 * pos + 1], g.xyz[3 * pos + 2]) <= r2Bound;
    if (__any_sync(0xffffffffu, hit)) {
      found = true;
      return true;
    }
    return false;
  };
  forCandidatesWarp<1>(g, cx, cy, cz, lane, test);
  if (!found)
    forCandidatesWarp<2>(g, cx, cy, cz, lane, test);
  return found;
}

__global__ void __launch_bounds__(kBlockThreads) tagEmptyKernel(const double *__restrict__ centers, int *__restrict__ tag, int64_t total, BinGridView g, double r2Bound)
{
  forLiveCentres(tag, total, [&](int64_t c, int lane) {
    const double cx = centers[3 * c], cy = centers[3 * c + 1], cz = centers[3 * c + 2];
    if (!anyVertexInside(g, cx, cy, cz, r2Bound, lane) && lane == 0)
      tag[c] = 1;
  });
}

// ---- tagEmptyClusters on the regular lattice, from the centre side (dense lattices) ----
// One warp per lattice centre and mesh; the same distance expression on the same operands as markLatticeKernel.
__global__ void __launch_bounds__(kBlockThreads) latticeEmptyKernel(LatticeAxes L, int64_t total, BinGridView g, double r2Bound, int *__restrict__ tag, int bit)
{
  const int64_t id   = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int     lane = threadIdx.x & 31;
  if (id >= total)
    return;
  const int    z = static_cast<int>(id % L.n[2]);
  const int    y = static_cast<int>((id / L.n[2]) % L.n[1]);
  const int    x = static_cast<int>(id / (static_cast<int64_t>(L.n[2]) * L.n[1]));
  if (!latticeLayerInRange(L, x, y, z))
    return; // another rank's part of the lattice
  const double cx = L.a[0][x], cy = L.a[1][y], cz = L.a[2][z];
  if (anyVertexInside(g, cx, cy, cz, r2Bound, lane) && lane == 0)
    atomicAnd(&tag[id], ~bit);
}

// ---- tagEmptyClusters on the regular lattice, from the vertex side ---------------------------
// Before the projection the centres are lattice points (x_i, y_j, z_k), so "the sphere of centre c holds a vertex" is
// decided from the vertices: every vertex clears the bit of the (2-3 per axis) lattice centres it is closer to than the
// radius. Work is proportional to the mesh, not to the lattice. The distance test is the one of tagEmptyKernel on the
// same operands (centre first), so the outcome is identical; the per-axis window is only a filter: |c_d - v_d| >= r
// implies dist >= r because IEEE multiplication, addition and square root are monotone.
__device__ __forceinline__ void latticeAxisRange(const double *__restrict__ ax, int n, double start, double invDist, double v, double radius, int &lo,
                                                 int &hi)
{
  if (n == 1) {
    lo = 0;
    hi = 0;
  } else {
    const double t = (v - start) * invDist, rho = radius * invDist;
    // the stored coordinates grow by repeated addition (a few ulp off start + i dist): one index of slack on each side
    lo = static_cast<int>(fmin(fmax(floor(t - rho) - 1.0, 0.0), static_cast<double>(n)));
    hi = static_cast<int>(fmin(fmax(ceil(t + rho) + 1.0, -1.0), static_cast<double>(n - 1)));
  }
  while (lo <= hi && !(fabs(__dsub_rn(ax[lo], v)) < radius))
    ++lo;
  while (hi >= lo && !(fabs(__dsub_rn(ax[hi], v)) < radius))
    --hi;
}

__global__ void __launch_bounds__(256) markLatticeKernel(const double *__restrict__ xyz, int64_t n, LatticeAxes L, double radius, double r2Bound, int *tag, int bit)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  const double vx = xyz[3 * i], vy = xyz[3 * i + 1], vz = xyz[3 * i + 2];
  int          x0, x1, y0, y1, z0, z1;
  latticeAxisRange(L.a[0], L.n[0], L.start[0], L.invDist[0], vx, radius, x0, x1);
  latticeAxisRange(L.a[1], L.n[1], L.start[1], L.invDist[1], vy, radius, y0, y1);
  latticeAxisRange(L.a[2], L.n[2], L.start[2], L.invDist[2], vz, radius, z0, z1);
  for (int x = x0; x <= x1; ++x) {
    const double cx = L.a[0][x];
    for (int y = y0; y <= y1; ++y) {
      const double cy = L.a[1][y];
      for (int z = z0; z <= z1; ++z) {
        if (!latticeLayerInRange(L, x, y, z))
          continue;
        const int64_t id = (static_cast<int64_t>(x) * L.n[1] + y) * L.n[2] + z; // impl/CreateClustering.hpp:36-45
        // most centres have long been cleared by another vertex (a stale read only costs a redundant test)
        if ((tag[id] & bit) && sqDist3(cx, cy, L.a[2][z], vx, vy, vz) <= r2Bound)
          atomicAnd(&tag[id], ~bit);
      }
    }
  }
}

// ---- projectClusterCentersToinputMesh: one warp per live centre ------------------------------
__global__ void __launch_bounds__(kBlockThreads) projectKernel(double *__restrict__ centers, const int *__restrict__ tag, int64_t total, BinGridView g)
{
  forLiveCentres(tag, total, [&](int64_t c, int lane) {
    const double cx = centers[3 * c], cy = centers[3 * c + 1], cz = centers[3 * c + 2];
    double       best = 1.7976931348623157e308;
    int          bestId = 0x7fffffff, bestPos = -1;
    auto visit = [&](int pos, bool valid) {
      if (valid) {
        const double d2 = sqDist3(cx, cy, cz, g.xyz[3 * pos], g.xyz[3 * pos + 1], g.xyz[3 * pos + 2]);
        const int    id = g.ids[pos];
        if (d2 < best || (d2 == best && id < bestId)) {
          best    = d2;
          bestId  = id;
          bestPos = pos;
        }
      }
      return false;
    };
    auto reduce = [&] {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int    oi = __shfl_xor_sync(0xffffffffu, bestId, o);
        const int    op = __shfl_xor_sync(0xffffffffu, bestPos, o);
        if (ob < best || (ob == best && oi < bestId)) {
          best    = ob;
          bestId  = oi;
          bestPos = op;
        }
      }
    };
    // The 3x3x3 cells around the centre first: every vertex outside of them is at least one cell edge away, so a vertex
    // found closer than that (with a margin for the rounding of the squared distances) is the nearest one of the mesh,
    // ties included (they lie inside the block as well). Otherwise the full 5x5x5 neighbourhood is searched.
    forCandidatesWarp<1>(g, cx, cy, cz, lane, visit);
    reduce();
    const double edge = 1.0 / fmax(g.ihx, fmax(g.ihy, g.ihz));
    if (!(best < edge * edge * (1.0 - 1e-9))) { // warp-uniform after the reduction
      best    = 1.7976931348623157e308;
      bestId  = 0x7fffffff;
      bestPos = -1;
      forCandidatesWarp<2>(g, cx, cy, cz, lane, visit);
      reduce();
    }
    if (lane < 3 && bestPos >= 0)
      centers[3 * c + lane] = g.xyz[3 * bestPos + lane];
  });
}

// ---- live list: the lattice centres that survive the two emptiness tests ----------------------
// Everything after tagEmptyLattice works on this list (ascending lattice index = the reference's visiting order); the
// full lattice only keeps its tag and the exclusive scan `pos` (lattice index -> list index) for neighbour look-ups.
__global__ void fillIntKernel(int *__restrict__ p, int64_t n, int v)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n)
    p[i] = v;
}
// tag0 for the lattice layers this rank processes, "empty" (all bits) for the others
__global__ void initLatticeTagKernel(int *__restrict__ tag, int64_t total, LatticeAxes L, int tag0)
{
  const int64_t id = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (id >= total)
    return;
  const int z = static_cast<int>(id % L.n[2]);
  const int y = static_cast<int>((id / L.n[2]) % L.n[1]);
  const int x = static_cast<int>(id / (static_cast<int64_t>(L.n[2]) * L.n[1]));
  tag[id]     = latticeLayerInRange(L, x, y, z) ? tag0 : 7;
}

__global__ void liveListKernel(int64_t total, const int *__restrict__ tag, const int *__restrict__ pos, LatticeAxes L, int *__restrict__ liveIds,
                               double *__restrict__ centers, int *__restrict__ liveTag)
{
  const int64_t id = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (id >= total || tag[id] != 0)
    return;
  // linear index (x * ny + y) * nz + z  (impl/CreateClustering.hpp:36-45)
  const int     z = static_cast<int>(id % L.n[2]);
  const int     y = static_cast<int>((id / L.n[2]) % L.n[1]);
  const int     x = static_cast<int>(id / (static_cast<int64_t>(L.n[2]) * L.n[1]));
  const int64_t o = pos[id];
  liveIds[o]         = static_cast<int>(id);
  centers[3 * o]     = L.a[0][x];
  centers[3 * o + 1] = L.a[1][y];
  centers[3 * o + 2] = L.a[2][z];
  liveTag[o]         = 0;
}

// ---- tagDuplicateCenters ---------------------------------------------------------------------
// state: 0 = untagged (final), 1 = tagged (final), 2 = undecided
__global__ void dupInitKernel(const double *__restrict__ centers, const int *__restrict__ liveIds, int64_t nLive, const int *__restrict__ tag,
                              const int *__restrict__ pos, int64_t total, NeighbourOffsets offs, double thr2, int *__restrict__ state,
                              unsigned *__restrict__ lowMask)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= nLive)
    return;
  const long long id = liveIds[i];
  const double    cx = centers[3 * i], cy = centers[3 * i + 1], cz = centers[3 * i + 2];
  bool            higher = false;
  unsigned        low    = 0;
  for (int k = 0; k < offs.n; ++k) {
    const long long off = offs.off[k];
    if (off == 0)
      continue; // CreateClustering.hpp:103-105
    const long long nb = id + off;
    if (nb < 0 || nb >= total || tag[nb] != 0)
      continue;
    const long long j = pos[nb];
    if (sqDist3(cx, cy, cz, centers[3 * j], centers[3 * j + 1], centers[3 * j + 2]) < thr2) {
      if (nb > id)
        higher = true;
      else
        low |= (1u << k);
    }
  }
  // a close neighbour with a HIGHER index has not been visited yet by the sequential sweep of the
  // reference, hence it is untagged when centre i is visited -> i gets tagged
  state[i]   = higher ? 1 : (low == 0 ? 0 : 2);
  lowMask[i] = low;
}

__global__ void dupResolveKernel(int64_t nLive, const int *__restrict__ liveIds, const int *__restrict__ pos, NeighbourOffsets offs, int *__restrict__ state,
                                 const unsigned *__restrict__ lowMask, int *__restrict__ undecided)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= nLive || state[i] != 2)
    return;
  const unsigned  low = lowMask[i];
  const long long id  = liveIds[i];
  bool            anyUntagged = false, anyUndecided = false;
  for (int k = 0; k < offs.n; ++k)
    if (low & (1u << k)) { // bit k set: that neighbour is live, pos[] is its list index
      const int s = reinterpret_cast<volatile int *>(state)[pos[id + offs.off[k]]];
      anyUntagged |= (s == 0);
      anyUndecided |= (s == 2);
    }
  if (anyUntagged)
    state[i] = 1;
  else if (!anyUndecided)
    state[i] = 0;
  else
    atomicAdd(undecided, 1);
}

__global__ void stateToTagKernel(int64_t total, const int *__restrict__ state, int *__restrict__ tag)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < total)
    tag[i] = state[i] != 0 ? 1 : 0;
}

// ---- removeTaggedVertices (stream compaction, order preserving) ------------------------------
__global__ void keepFlagKernel(int64_t total, const int *__restrict__ tag, int *__restrict__ keep)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < total)
    keep[i] = tag[i] == 0 ? 1 : 0;
}
__global__ void compactKernel(int64_t total, const int *__restrict__ tag, const int *__restrict__ pos, const double *__restrict__ centers, double *__restrict__ out)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < total && tag[i] == 0) {
    const int64_t o = pos[i];
    out[3 * o]      = centers[3 * i];
    out[3 * o + 1]  = centers[3 * i + 1];
    out[3 * o + 2]  = centers[3 * i + 2];
  }
}

// ---- cluster membership: one warp per cluster ------------------------------------------------
// warp-cooperative ascending sort of seg[0..n) (global memory) through a shared staging buffer
__device__ void warpSortSegment(int *seg, int n, int *stage, int stageCap, int lane)
{
  if (n <= 1)
    return;
  if (n <= stageCap) {
    int p = 1;
    while (p < n)
      p <<= 1;
    for (int i = lane; i < p; i += 32)
      stage[i] = i < n ? seg[i] : 0x7fffffff;
    __syncwarp();
    for (int k = 2; k <= p; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < p; i += 32) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const int  a = stage[i], b = stage[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) {
              stage[i]   = b;
              stage[ixj] = a;
            }
          }
        }
        __syncwarp();
      }
    for (int i = lane; i < n; i += 32)
      seg[i] = stage[i];
    __syncwarp();
  } else {
    // rare: very large cluster — odd-even transposition sort in global memory
    for (int round = 0; round < n; ++round) {
      for (int i = 2 * lane + (round & 1); i + 1 < n; i += 64) {
        const int a = seg[i], b = seg[i + 1];
        if (a > b) {
          seg[i]     = b;
          seg[i + 1] = a;
        }
      }
      __syncwarp();
    }
  }
}

constexpr int kSortCap = 1024; // ints staged per warp

// slab filter of the sharded mapping (pum.cuh, PumSlab): the vertex at binned position pos belongs to this rank
__device__ __forceinline__ bool inSlab(const BinGridView &g, int pos, const PumSlab &slab)
{
  const double t = g.xyz[3 * static_cast<long long>(pos) + slab.axis];
  return t >= slab.lo && t < slab.hi;
}

__device__ void gatherMembers(const BinGridView &g, double cx, double cy, double cz, double r2Bound, int *seg, int lane, const PumSlab *slab = nullptr)
{
  int base = 0;
  forCandidatesWarp<2>(g, cx, cy, cz, lane, [&](int pos, bool valid) {
    if (slab && valid)
      valid = inSlab(g, pos, *slab);
    const bool     hit = valid && sqDist3(cx, cy, cz, g.xyz[3 * pos], g.xyz[3 * pos + 1], g.xyz[3 * pos + 2]) <= r2Bound;
    const unsigned m   = __ballot_sync(0xffffffffu, hit);
    if (hit)
      seg[base + __popc(m & ((1u << lane) - 1u))] = g.ids[pos];
    base += __popc(m);
    return false;
  });
  __syncwarp();
}

// ---- single-pass membership: candidates are visited once; the sorted vertex sets go to fixed-size slots, sizes are
// scanned, and a copy kernel packs the slots into the CSR arrays. Clusters that do not fit a slot raise `overflow`;
// the host then falls back to the two-pass fill above (the sizes are exact in either case). ----
constexpr int kSlotCap = 128;

// ids of the vertices of grid g within radius r of the centre -> stage (unsorted, first kSlotCap hits); returns the count.
// WEIGHTS (out-mesh pass): every vertex within rW of the centre also receives this cluster's partition-of-unity weight,
// PartitionOfUnityMapping.hpp:306-333 seen from the cluster side: w_c = C2(|x - c| / r) is added to the weight sum of
// the vertex and its cluster count is incremented (the reference loops over the clusters of a vertex instead).
template <bool WEIGHTS, bool SLAB = false>
__device__ __forceinline__ int gatherToShared(const BinGridView &g, double cx, double cy, double cz, double r2Bound, int *stage, int *stagePos, int lane,
                                              double rW2Bound = 0.0, double rinv = 0.0, double *__restrict__ wsum = nullptr, int *__restrict__ wcount = nullptr,
                                              const PumSlab *slab = nullptr)
{
  int base = 0;
  forCandidatesWarp<2>(g, cx, cy, cz, lane, [&](int pos, bool valid) {
    double d2 = 1.7976931348623157e308;
    if (SLAB && valid) // sharded mapping: only the output vertices this rank owns are listed and weighted
      valid = inSlab(g, pos, *slab);
    if (valid) // computeWeight (SphericalVertexCluster.hpp:337-344): squared difference (centre, vertex)
      d2 = sqDist3(cx, cy, cz, g.xyz[3 * pos], g.xyz[3 * pos + 1], g.xyz[3 * pos + 2]);
    const bool     hit = d2 <= r2Bound; // sqrt(d2) < r, see strictRadiusBound
    const unsigned m   = __ballot_sync(0xffffffffu, hit);
    const int      at  = base + __popc(m & ((1u << lane) - 1u));
    int            id  = 0;
    if (hit || (WEIGHTS && d2 <= rW2Bound))
      id = g.ids[pos];
    if (hit && at < kSlotCap) {
      stage[at]    = id;
      stagePos[at] = pos; // position in the binned arrays: neighbouring hits share cache lines there
    }
    if (WEIGHTS && d2 <= rW2Bound) {
      // wcount only counts the coverings with weight exactly zero (a vertex within an ulp of the cluster radius): it is
      // what tells "covered, all weights zero -> equal weights" (PartitionOfUnityMapping.hpp:329-333) from "not covered"
      // (:314-318); one atomic per covering instead of two
      const double wgt = pumWeight(__dsqrt_rn(d2), rinv);
      if (wgt > 0.0)
        atomicAdd(wsum + id, wgt);
      else
        atomicAdd(wcount + id, 1);
    }
    base += __popc(m);
    return false;
  });
  __syncwarp();
  return base;
}

// ascending rank sort of the n <= kSlotCap distinct ids in stage; dst[rank] = payload of that id (its binned position)
// 1 if a < b, for vertex IDs (0 <= a, b < 2^30): the sign bit of the difference; ptxas turns the comparison form into a
// three-instruction select per element, this is a subtraction and a shift-add
__device__ __forceinline__ int lessThan(int a, int b) { return static_cast<int>((static_cast<unsigned>(a) - static_cast<unsigned>(b)) >> 31); }

__device__ __forceinline__ void rankSortToGlobal(const int *stage, const int *payload, int n, int *__restrict__ dst, int lane)
{
  // ranks by counting: four independent counters per element (one compare + one predicated increment each; a single
  // counter compiled to a three-instruction dependent chain per compare), two elements per lane share the loads
  for (int i = lane; i < n; i += 64) {
    const int  v = stage[i];
    const bool two = i + 32 < n;
    const int  w = two ? stage[i + 32] : 0;
    int        a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    int        j = 0;
    if (__any_sync(__activemask(), two)) {
      for (; j + 4 <= n; j += 4) {
        const int4 q = *reinterpret_cast<const int4 *>(stage + j);
        a0 += lessThan(q.x, v);
        a1 += lessThan(q.y, v);
        a2 += lessThan(q.z, v);
        a3 += lessThan(q.w, v);
        b0 += lessThan(q.x, w);
        b1 += lessThan(q.y, w);
        b2 += lessThan(q.z, w);
        b3 += lessThan(q.w, w);
      }
    } else {
      for (; j + 4 <= n; j += 4) {
        const int4 q = *reinterpret_cast<const int4 *>(stage + j);
        a0 += lessThan(q.x, v);
        a1 += lessThan(q.y, v);
        a2 += lessThan(q.z, v);
        a3 += lessThan(q.w, v);
      }
    }
    for (; j < n; ++j) {
      const int q = stage[j];
      a0 += lessThan(q, v);
      b0 += lessThan(q, w);
    }
    dst[(a0 + a1) + (a2 + a3)] = payload[i];
    if (two)
      dst[(b0 + b1) + (b2 + b3)] = payload[i + 32];
  }
}

template <bool SLAB>
__global__ void memberGatherKernel(const double *__restrict__ centers, int64_t nClusters, BinGridView gin, BinGridView gout, double rIn, double rInBound,
                                   double rOutBound,
                                   int *__restrict__ nIn, int *__restrict__ nOut, long long *__restrict__ tri, int *__restrict__ slotIn,
                                   int *__restrict__ slotOut, int *__restrict__ overflow, double *__restrict__ wsum, int *__restrict__ wcount, PumSlab slab,
                                   bool fullMatrices)
{
  __shared__ __align__(16) int stage[kWarpsPerBlock][kSlotCap];
  __shared__ int               stagePos[kWarpsPerBlock][kSlotCap];
  const int64_t c    = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int     lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (c >= nClusters)
    return;
  const double cx = centers[3 * c], cy = centers[3 * c + 1], cz = centers[3 * c + 2];
  const int    ci = gatherToShared<false>(gin, cx, cy, cz, rInBound, stage[w], stagePos[w], lane);
  if (ci <= kSlotCap)
    rankSortToGlobal(stage[w], stagePos[w], ci, slotIn + c * kSlotCap, lane);
  __syncwarp();
  const int co = gatherToShared<true, SLAB>(gout, cx, cy, cz, rOutBound, stage[w], stagePos[w], lane, rInBound, 1.0 / rIn, wsum, wcount, &slab);
  if (co <= kSlotCap)
    rankSortToGlobal(stage[w], stagePos[w], co, slotOut + c * kSlotCap, lane);
  if (lane == 0) {
    nIn[c]  = ci;
    nOut[c] = co;
    tri[c]  = pumMatSize(ci, fullMatrices); // even count: 16-byte aligned factors for the TMA bulk copy
    if (ci > kSlotCap || co > kSlotCap)
      atomicOr(overflow, 1);
  }
}

// Also writes the cluster-ordered records the algebra kernels read: in-vertex coordinates and, per out-vertex, coordinates
// + normalised partition-of-unity weight (the weight sums are final once the gather kernel has finished).
__global__ void memberPackKernel(int64_t nClusters, const int *__restrict__ nIn, const int *__restrict__ nOut, const long long *__restrict__ inOff,
                                 const long long *__restrict__ outOff, const int *__restrict__ slotIn, const int *__restrict__ slotOut,
                                 int *__restrict__ inIds, int *__restrict__ outIds, const double *__restrict__ centers, double radius,
                                 BinGridView gin, BinGridView gout, const double *__restrict__ wsum,
                                 const int *__restrict__ wcount, double *__restrict__ inRec, double4 *__restrict__ outRec)
{
  const int64_t c    = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int     lane = threadIdx.x & 31;
  if (c >= nClusters)
    return;
  const int  ni = nIn[c], no = nOut[c];
  int       *di = inIds + inOff[c], *dO = outIds + outOff[c];
  const int *si = slotIn + c * kSlotCap, *so = slotOut + c * kSlotCap;
  const double cx = centers[3 * c], cy = centers[3 * c + 1], cz = centers[3 * c + 2];
  const double rinv = 1.0 / radius;
  double      *ri = inRec + 3 * inOff[c];
  double4     *ro = outRec + outOff[c];
  for (int i = lane; i < ni; i += 32) { // the slots hold positions in the binned arrays, in ascending-ID order
    const int pos = si[i];
    di[i]         = gin.ids[pos];
    const double *p = gin.xyz + 3 * static_cast<long long>(pos);
    ri[3 * i]       = p[0];
    ri[3 * i + 1]   = p[1];
    ri[3 * i + 2]   = p[2];
  }
  for (int i = lane; i < no; i += 32) {
    const int pos = so[i];
    const int id  = gout.ids[pos];
    dO[i]         = id;
    const double *p = gout.xyz + 3 * static_cast<long long>(pos);
    const double  x = p[0], y = p[1], z = p[2];
    ro[i]           = make_double4(x, y, z, normalizedWeight(x, y, z, cx, cy, cz, rinv, wsum[id], wcount + id));
  }
}

__global__ void memberFillKernel(const double *__restrict__ centers, int64_t nClusters, BinGridView gin, BinGridView gout, double rInBound, double rOutBound,
                                 const long long *__restrict__ inOff, const long long *__restrict__ outOff, int *__restrict__ inIds, int *__restrict__ outIds,
                                 PumSlab slab, bool useSlab)
{
  __shared__ int stage[kWarpsPerBlock][kSortCap];
  const int64_t  c    = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (c >= nClusters)
    return;
  const double cx = centers[3 * c], cy = centers[3 * c + 1], cz = centers[3 * c + 2];
  int         *segIn = inIds + inOff[c], *segOut = outIds + outOff[c];
  const int    ni = static_cast<int>(inOff[c + 1] - inOff[c]), no = static_cast<int>(outOff[c + 1] - outOff[c]);
  gatherMembers(gin, cx, cy, cz, rInBound, segIn, lane);
  warpSortSegment(segIn, ni, stage[w], kSortCap, lane);
  gatherMembers(gout, cx, cy, cz, rOutBound, segOut, lane, useSlab ? &slab : nullptr);
  warpSortSegment(segOut, no, stage[w], kSortCap, lane);
}

// ---- out-vertices that no cluster covers (PartitionOfUnityMapping.hpp:312-320) ----
__global__ void unassignedKernel(const double *__restrict__ wsum, const int *__restrict__ wcount, int64_t nOut, int *__restrict__ firstUnassigned,
                                 const double *__restrict__ outXyz,
                                 PumSlab slab, bool useSlab)
{
  const int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (v < nOut && !(wsum[v] > 0.0) && wcount[v] == 0) { // no covering cluster at all
    if (useSlab) { // sharded mapping: only the vertices this rank owns have complete counts
      const double t = outXyz[3 * v + slab.axis];
      if (!(t >= slab.lo && t < slab.hi))
        return;
    }
    atomicMin(firstUnassigned, static_cast<int>(v));
  }
}

// ---- sharded mapping: slab decomposition of the (global) centre list ----
__global__ void vertexHistogramKernel(const double *__restrict__ xyz, int64_t n, int axis, double lo, double invBin, int nBins, int *__restrict__ hist)
{
  __shared__ int local[4096];
  for (int b = threadIdx.x; b < nBins; b += blockDim.x)
    local[b] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double t = floor((xyz[3 * i + axis] - lo) * invBin);
    atomicAdd(&local[static_cast<int>(fmin(fmax(t, 0.0), static_cast<double>(nBins - 1)))], 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nBins; b += blockDim.x)
    if (local[b])
      atomicAdd(&hist[b], local[b]);
}

__global__ void slabFlagKernel(const double *__restrict__ centers, int64_t n, int axis, double lo, double hi, int *__restrict__ keep)
{
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c < n) {
    const double t = centers[3 * c + axis];
    keep[c]        = (t >= lo && t < hi) ? 1 : 0;
  }
}

__global__ void slabCompactKernel(int64_t n, const int *__restrict__ keep, const int *__restrict__ pos, const double *__restrict__ centers,
                                  double *__restrict__ out, int *__restrict__ globalIdx)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n && keep[i]) {
    const int64_t o = pos[i];
    out[3 * o]      = centers[3 * i];
    out[3 * o + 1]  = centers[3 * i + 1];
    out[3 * o + 2]  = centers[3 * i + 2];
    globalIdx[o]    = static_cast<int>(i);
  }
}

// ---- size buckets + statistics ----------------------------------------------------------------
// acc: [0..12] bucket counts, [13] sum n, [14] sum m, [15] sum n^2, [16] sum n^3, [17] sum m n, [18] max n, [19..31] max n per bucket
// sizeHist: clusters per size class (pumSizeClassOf)
__global__ void classifyKernel(const int *__restrict__ nIn, const int *__restrict__ nOut, int64_t nClusters, unsigned long long *__restrict__ acc,
                               int *__restrict__ sizeHist)
{
  // Everything is collected per block in shared memory first: the global counters are a few dozen addresses that every
  // block hits (one atomic per warp and counter made this kernel 0.16 ms at 0.94 M clusters).
  __shared__ int                localHist[kNumSizeClasses];
  __shared__ unsigned long long localAcc[32]; // same slots as `acc`: [b] bucket counts, [13..17] sums, [18] max n, [19 + b] bucket max n
  for (int k = threadIdx.x; k < kNumSizeClasses; k += blockDim.x)
    localHist[k] = 0;
  if (threadIdx.x < 32)
    localAcc[threadIdx.x] = 0;
  __syncthreads();
  const int64_t c    = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int     lane = threadIdx.x & 31;
  unsigned long long n = 0, m = 0;
  int                b = -1;
  if (c < nClusters) {
    n = static_cast<unsigned long long>(nIn[c]);
    m = static_cast<unsigned long long>(nOut[c]);
    b = pumBucketOf(static_cast<int>(n));
    atomicAdd(&localHist[pumSizeClassOf(static_cast<int>(n))], 1);
  }
  unsigned long long v[5] = {n, m, n * n, n * n * n, m * n};
  unsigned long long mx   = n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 5; ++q)
      v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 5; ++q)
      atomicAdd(&localAcc[13 + q], v[q]);
    atomicMax(&localAcc[18], mx);
  }
  // warp-aggregated bucket counters
  const unsigned peers = __match_any_sync(0xffffffffu, b);
  if (b >= 0) {
    unsigned long long bm = n;
    for (unsigned rest = peers; rest;) { // max n among the peers of this bucket
      const int src = __ffs(rest) - 1;
      rest &= rest - 1;
      bm = max(bm, __shfl_sync(peers, n, src));
    }
    if (lane == __ffs(peers) - 1) {
      atomicAdd(&localAcc[b], static_cast<unsigned long long>(__popc(peers)));
      atomicMax(&localAcc[19 + b], bm);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kNumSizeClasses; k += blockDim.x)
    if (localHist[k])
      atomicAdd(&sizeHist[k], localHist[k]);
  if (threadIdx.x < 32 && localAcc[threadIdx.x]) {
    const int k = threadIdx.x;
    if (k == 18 || k >= 19)
      atomicMax(&acc[k], localAcc[k]);
    else
      atomicAdd(&acc[k], localAcc[k]);
  }
}

__global__ void bucketScatterKernel(const int *__restrict__ nIn, const int *__restrict__ nOut, const long long *__restrict__ inOff,
                                    const long long *__restrict__ outOff, const long long *__restrict__ matOff, int64_t nClusters,
                                    const long long *__restrict__ start, int *__restrict__ cursor, int *__restrict__ order, PumEntry *__restrict__ entries)
{
  const int64_t c    = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int     lane = threadIdx.x & 31;
  const int     b    = c < nClusters ? pumSizeClassOf(nIn[c]) : -1; // size class: the entries end up sorted by size
  const unsigned peers = __match_any_sync(0xffffffffu, b);
  if (b >= 0) {
    const int leader = __ffs(peers) - 1;
    int       base   = 0;
    if (lane == leader)
      base = atomicAdd(&cursor[b], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const long long at = start[b] + base + __popc(peers & ((1u << lane) - 1u));
    order[at]          = static_cast<int>(c);
    entries[at]        = PumEntry{static_cast<int>(c), nIn[c], nOut[c], 0, inOff[c], outOff[c], matOff[c], 0};
  }
}

} // namespace

// =============================================================================================
void PumClustering::initLattice(const LatticeAxes &ax, int tag0, cudaStream_t s)
{
  axes  = ax;
  total = static_cast<int64_t>(ax.n[0]) * ax.n[1] * ax.n[2];
  nLive = 0;
  tag.alloc(total);
  if (ax.rangeAxis >= 0)
    initLatticeTagKernel<<<divUp(total, 256), 256, 0, s>>>(tag.p, total, ax, tag0);
  else
    fillIntKernel<<<divUp(total, 256), 256, 0, s>>>(tag.p, total, tag0);
  RBF_LAUNCHED();
}

void PumClustering::tagEmptyLattice(const BinGrid &mesh, int bit, cudaStream_t s)
{
  if (mesh.n == 0)
    return; // every centre keeps its bit = tagged
  // dense lattice (volume meshes: a centre per ~vertices-per-cluster / 5 vertices): one warp per centre, which finds a
  // vertex in its first batch of candidates; sparse lattice (surface meshes in 3-D: 97 % of the lattice is empty): from
  // the vertex side, work proportional to the mesh. Both decide every centre with the same distance test.
  int64_t active = total; // lattice centres this rank processes
  if (axes.rangeAxis >= 0)
    active = total / std::max(1, axes.n[axes.rangeAxis]) * std::max(0, axes.rangeHi - axes.rangeLo + 1);
  if (active * 8 <= mesh.n)
    latticeEmptyKernel<<<warpGrid(total), kBlockThreads, 0, s>>>(axes, total, mesh.view(), strictRadiusBound(radius), tag.p, bit);
  else
    markLatticeKernel<<<divUp(mesh.n, 256), 256, 0, s>>>(mesh.xyz.p, mesh.n, axes, radius, strictRadiusBound(radius), tag.p, bit);
  RBF_LAUNCHED();
}

int64_t PumClustering::gatherLive(cudaStream_t s)
{
  DevBuf<int> keep;
  keep.alloc(total);
  pos.alloc(total + 1);
  const int grid = divUp(total, 256);
  keepFlagKernel<<<grid, 256, 0, s>>>(total, tag.p, keep.p);
  RBF_LAUNCHED();
  exclusiveScan(keep.p, pos.p, total, s);
  int count = 0;
  RBF_CUDA(cudaMemcpyAsync(&count, pos.p + total, sizeof(int), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  nLive = count;
  liveIds.alloc(std::max(count, 1));
  liveTag.alloc(std::max(count, 1));
  centers.alloc(3 * static_cast<size_t>(std::max(count, 1)));
  if (count > 0) {
    liveListKernel<<<grid, 256, 0, s>>>(total, tag.p, pos.p, axes, liveIds.p, centers.p, liveTag.p);
    RBF_LAUNCHED();
  }
  return nLive;
}

void PumClustering::tagEmpty(const BinGrid &mesh, cudaStream_t s)
{
  if (nLive == 0)
    return;
  tagEmptyKernel<<<divUp(nLive, kBlockThreads), kBlockThreads, 0, s>>>(centers.p, liveTag.p, nLive, mesh.view(), strictRadiusBound(radius));
  RBF_LAUNCHED();
}

void PumClustering::project(const BinGrid &inMesh, cudaStream_t s)
{
  if (nLive == 0)
    return;
  projectKernel<<<divUp(nLive, kBlockThreads), kBlockThreads, 0, s>>>(centers.p, liveTag.p, nLive, inMesh.view());
  RBF_LAUNCHED();
}

void PumClustering::tagDuplicates(const NeighbourOffsets &offs, double threshold, cudaStream_t s)
{
  if (nLive == 0)
    return;
  DevBuf<int>      state, undecided;
  DevBuf<unsigned> low;
  state.alloc(nLive);
  low.alloc(nLive);
  undecided.alloc(1);
  const int grid = divUp(nLive, 256);
  // CreateClustering.hpp:112: squared distance < pow_int<2>(radius)
  dupInitKernel<<<grid, 256, 0, s>>>(centers.p, liveIds.p, nLive, tag.p, pos.p, total, offs, threshold * threshold, state.p, low.p);
  RBF_LAUNCHED();
  for (int64_t it = 0; it <= nLive; ++it) {
    RBF_CUDA(cudaMemsetAsync(undecided.p, 0, sizeof(int), s));
    dupResolveKernel<<<grid, 256, 0, s>>>(nLive, liveIds.p, pos.p, offs, state.p, low.p, undecided.p);
    RBF_LAUNCHED();
    int left = 0;
    RBF_CUDA(cudaMemcpyAsync(&left, undecided.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    RBF_CUDA(cudaStreamSynchronize(s));
    if (left == 0)
      break;
  }
  stateToTagKernel<<<grid, 256, 0, s>>>(nLive, state.p, liveTag.p);
  RBF_LAUNCHED();
}

int64_t PumClustering::compact(DevBuf<double> &out, cudaStream_t s)
{
  if (nLive == 0) {
    out.alloc(3);
    return 0;
  }
  DevBuf<int> keep, opos;
  keep.alloc(nLive);
  opos.alloc(nLive + 1);
  const int grid = divUp(nLive, 256);
  keepFlagKernel<<<grid, 256, 0, s>>>(nLive, liveTag.p, keep.p);
  RBF_LAUNCHED();
  exclusiveScan(keep.p, opos.p, nLive, s);
  int count = 0;
  RBF_CUDA(cudaMemcpyAsync(&count, opos.p + nLive, sizeof(int), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  out.alloc(3 * static_cast<size_t>(std::max(count, 1)));
  if (count > 0) {
    compactKernel<<<grid, 256, 0, s>>>(nLive, liveTag.p, opos.p, centers.p, out.p);
    RBF_LAUNCHED();
  }
  return count;
}

void pumBuildMembership(const double *dCenters, int64_t nClusters, const BinGrid &inMesh, const BinGrid &outMesh, double radius, PumMembership &m,
                        PumScratch &scratch, double *wsum, int *wcount, const double *inXyz, const double *outXyz, DevBuf<double> &inRec,
                        DevBuf<double4> &outRec, cudaStream_t s, const PumSlab *outSlab, const std::function<void()> &weightsComplete, bool fullMatrices)
{
  const PumSlab slab = outSlab ? *outSlab : PumSlab{};
  const double rOut = radius - kZeroDiff; // SphericalVertexCluster.hpp:155
  const double rInBound = strictRadiusBound(radius), rOutBound = strictRadiusBound(rOut);
  DevBuf<int>  nIn, nOut;
  nIn.alloc(nClusters);
  nOut.alloc(nClusters);
  m.tri.alloc(nClusters);
  m.inOff.alloc(nClusters + 1);
  m.outOff.alloc(nClusters + 1);
  m.matOff.alloc(nClusters + 1);
  m.nIn.alloc(nClusters);
  m.nOut.alloc(nClusters);
  ScratchBuf<int> &slotIn = scratch.slotIn, &slotOut = scratch.slotOut;
  const size_t     slots  = static_cast<size_t>(std::max<int64_t>(nClusters, 1)) * kSlotCap;
  if (slotIn.n < slots) { // grow only
    RBF_CUDA(cudaStreamSynchronize(s));
    slotIn.reserve(slots + slots / 8);
    slotOut.reserve(slots + slots / 8);
  }
  DevBuf<int> overflow;
  overflow.alloc(1);
  RBF_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
  if (outSlab)
    memberGatherKernel<true><<<warpGrid(nClusters), kBlockThreads, 0, s>>>(dCenters, nClusters, inMesh.view(), outMesh.view(), radius, rInBound, rOutBound, m.nIn.p, m.nOut.p,
                                                                          reinterpret_cast<long long *>(m.tri.p), slotIn.p, slotOut.p, overflow.p, wsum, wcount, slab, fullMatrices);
  else
    memberGatherKernel<false><<<warpGrid(nClusters), kBlockThreads, 0, s>>>(dCenters, nClusters, inMesh.view(), outMesh.view(), radius, rInBound, rOutBound, m.nIn.p, m.nOut.p,
                                                                           reinterpret_cast<long long *>(m.tri.p), slotIn.p, slotOut.p, overflow.p, wsum, wcount, slab, fullMatrices);
  RBF_LAUNCHED();
  if (weightsComplete) // sharded mapping, exchange mode: the weight sums of the other ranks' clusters are added here
    weightsComplete();
  exclusiveScan(m.nIn.p, m.inOff.p, nClusters, s);
  exclusiveScan(m.nOut.p, m.outOff.p, nClusters, s);
  exclusiveScan(m.tri.p, m.matOff.p, nClusters, s);
  int64_t totals[3];
  int     over = 0;
  RBF_CUDA(cudaMemcpyAsync(&totals[0], m.inOff.p + nClusters, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaMemcpyAsync(&totals[1], m.outOff.p + nClusters, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaMemcpyAsync(&totals[2], m.matOff.p + nClusters, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaMemcpyAsync(&over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  m.sumIn  = totals[0];
  m.sumOut = totals[1];
  m.sumTri = totals[2];
  m.inIds.alloc(std::max<int64_t>(m.sumIn, 1));
  m.outIds.alloc(std::max<int64_t>(m.sumOut, 1));
  inRec.alloc(3 * std::max<int64_t>(m.sumIn, 1));
  outRec.alloc(std::max<int64_t>(m.sumOut, 1));
  if (!over) {
    memberPackKernel<<<warpGrid(nClusters), kBlockThreads, 0, s>>>(nClusters, m.nIn.p, m.nOut.p, reinterpret_cast<const long long *>(m.inOff.p),
                                                                  reinterpret_cast<const long long *>(m.outOff.p), slotIn.p, slotOut.p, m.inIds.p, m.outIds.p,
                                                                  dCenters, radius, inMesh.view(), outMesh.view(), wsum, wcount, inRec.p, outRec.p);
    RBF_LAUNCHED();
  } else { // clusters beyond the slot size: two-pass fill with the (exact) sizes counted above, then the records
    memberFillKernel<<<warpGrid(nClusters), kBlockThreads, 0, s>>>(dCenters, nClusters, inMesh.view(), outMesh.view(), rInBound, rOutBound,
                                                                  reinterpret_cast<const long long *>(m.inOff.p), reinterpret_cast<const long long *>(m.outOff.p),
                                                                  m.inIds.p, m.outIds.p, slab, outSlab != nullptr);
    RBF_LAUNCHED();
    PumSolveArgs a{};
    a.radius = radius;
    a.inXyz = inXyz, a.outXyz = outXyz, a.centers = dCenters;
    a.nOut = m.nOut.p, a.outOff = reinterpret_cast<const long long *>(m.outOff.p);
    a.inIds = m.inIds.p, a.outIds = m.outIds.p;
    a.wsum = wsum, a.wcount = wcount;
    pumBuildInRecords(a, m.sumIn, inRec.p, s);
    pumBuildOutRecords(a, nClusters, outRec.p, s);
  }
  RBF_CUDA(cudaGetLastError());
}

// ---- membership of arbitrary points (just-in-time read mapping): the points within the cluster radius of every centre,
// sorted ascending per cluster, plus their partition-of-unity weight sums; PartitionOfUnityMapping.hpp:290-341 seen
// from the cluster side. Returns false if a cluster holds more than kSlotCap points (the caller falls back). ----
__global__ void memberGatherPointsKernel(const double *__restrict__ centers, int64_t nClusters, BinGridView gpts, double radius, double rBound, int *__restrict__ nPts,
                                         int *__restrict__ slot, int *__restrict__ overflow, double *__restrict__ wsum, int *__restrict__ wcount)
{
  __shared__ __align__(16) int stage[kWarpsPerBlock][kSlotCap];
  const int64_t c    = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int     lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (c >= nClusters)
    return;
  const double cx = centers[3 * c], cy = centers[3 * c + 1], cz = centers[3 * c + 2];
  __shared__ int stagePos[kWarpsPerBlock][kSlotCap];
  const int      co = gatherToShared<true>(gpts, cx, cy, cz, rBound, stage[w], stagePos[w], lane, rBound, 1.0 / radius, wsum, wcount);
  if (co <= kSlotCap)
    rankSortToGlobal(stage[w], stage[w], co, slot + c * kSlotCap, lane); // payload = the point ID itself
  if (lane == 0) {
    nPts[c] = co;
    if (co > kSlotCap)
      atomicOr(overflow, 1);
  }
}

__global__ void memberPackPointsKernel(int64_t nClusters, const int *__restrict__ nPts, const long long *__restrict__ off, const int *__restrict__ slot,
                                       int *__restrict__ ids)
{
  const int64_t c    = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int     lane = threadIdx.x & 31;
  if (c >= nClusters)
    return;
  const int  n = nPts[c];
  int       *d = ids + off[c];
  const int *q = slot + c * kSlotCap;
  for (int i = lane; i < n; i += 32)
    d[i] = q[i];
}

bool pumBuildPointMembership(const double *dCenters, int64_t nClusters, const BinGrid &points, double radius, PumScratch &scratch, DevBuf<int> &nPts,
                             DevBuf<int64_t> &off, DevBuf<int> &ids, double *wsum, int *wcount, cudaStream_t s)
{
  nPts.alloc(nClusters);
  off.alloc(nClusters + 1);
  const size_t slots = static_cast<size_t>(std::max<int64_t>(nClusters, 1)) * kSlotCap;
  if (scratch.slotOut.n < slots) {
    RBF_CUDA(cudaStreamSynchronize(s));
    scratch.slotOut.reserve(slots + slots / 8);
  }
  DevBuf<int> overflow;
  overflow.alloc(1);
  RBF_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
  memberGatherPointsKernel<<<warpGrid(nClusters), kBlockThreads, 0, s>>>(dCenters, nClusters, points.view(), radius, strictRadiusBound(radius), nPts.p, scratch.slotOut.p, overflow.p, wsum,
                                                                        wcount);
  RBF_LAUNCHED();
  exclusiveScan(nPts.p, off.p, nClusters, s);
  int64_t total = 0;
  int     over  = 0;
  RBF_CUDA(cudaMemcpyAsync(&total, off.p + nClusters, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaMemcpyAsync(&over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  if (over)
    return false;
  ids.alloc(std::max<int64_t>(total, 1));
  memberPackPointsKernel<<<warpGrid(nClusters), kBlockThreads, 0, s>>>(nClusters, nPts.p, reinterpret_cast<const long long *>(off.p), scratch.slotOut.p, ids.p);
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
  return true;
}

// ---- cluster-ordered out-vertex records: coordinates + normalised weight (PartitionOfUnityMapping.hpp:321-336), static per mapping ----
__global__ void outRecordKernel(PumSolveArgs a, int64_t nClusters, double4 *__restrict__ rec)
{
  const int64_t c    = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int     lane = threadIdx.x & 31;
  if (c >= nClusters)
    return;
  const int       m   = a.nOut[c];
  const long long off = a.outOff[c];
  const double    cx = a.centers[3 * c], cy = a.centers[3 * c + 1], cz = a.centers[3 * c + 2];
  const double    rinv = 1.0 / a.radius;
  for (int o = lane; o < m; o += 32) {
    const long long gid = a.outIds[off + o];
    const double   *p   = a.outXyz + 3 * gid;
    const double    x = p[0], y = p[1], z = p[2];
    rec[off + o]        = make_double4(x, y, z, normalizedWeight(x, y, z, cx, cy, cz, rinv, a.wsum[gid], a.wcount + gid));
  }
}

void pumBuildOutRecords(const PumSolveArgs &a, int64_t nClusters, double4 *rec, cudaStream_t s)
{
  if (nClusters <= 0)
    return;
  outRecordKernel<<<warpGrid(nClusters), kBlockThreads, 0, s>>>(a, nClusters, rec);
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

// ---- cluster-ordered in-vertex coordinates: one thread per CSR entry ----
__global__ void inRecordKernel(const int *__restrict__ inIds, const double *__restrict__ inXyz, int64_t sumIn, double *__restrict__ rec)
{
  const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (e >= sumIn)
    return;
  const double *p = inXyz + 3 * static_cast<long long>(inIds[e]);
  rec[3 * e]      = p[0];
  rec[3 * e + 1]  = p[1];
  rec[3 * e + 2]  = p[2];
}

void pumBuildInRecords(const PumSolveArgs &a, int64_t sumIn, double *rec, cudaStream_t s)
{
  if (sumIn <= 0)
    return;
  inRecordKernel<<<divUp(sumIn, 256), 256, 0, s>>>(a.inIds, a.inXyz, sumIn, rec);
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

void pumMarkFirstUnassigned(const double *wsum, const int *wcount, int64_t nOut, cudaStream_t s, const double *outXyz, const PumSlab *slab, int *dFirst)
{
  if (nOut > 0) {
    unassignedKernel<<<divUp(nOut, 256), 256, 0, s>>>(wsum, wcount, nOut, dFirst, outXyz, slab ? *slab : PumSlab{}, slab != nullptr);
    RBF_LAUNCHED();
  }
}

int64_t pumFirstUnassigned(const double *wsum, const int *wcount, int64_t nOut, cudaStream_t s, const double *outXyz, const PumSlab *slab,
                           const ShardGroup *group)
{
  DevBuf<int> first;
  first.alloc(1);
  RBF_CUDA(cudaMemsetAsync(first.p, 0x7f, sizeof(int), s));
  if (nOut > 0) {
    unassignedKernel<<<divUp(nOut, 256), 256, 0, s>>>(wsum, wcount, nOut, first.p, outXyz, slab ? *slab : PumSlab{}, slab != nullptr);
    RBF_LAUNCHED();
  }
  if (group) // every rank reports the same (lowest) unassigned vertex, so that all ranks raise the same error
    group->allReduceMin(first.p, 1, s);
  int f = 0;
  RBF_CUDA(cudaMemcpyAsync(&f, first.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  return f == 0x7f7f7f7f ? -1 : f;
}

} // namespace rbfmap

namespace rbfmap {
// ---- deterministic blend: vertex -> CSR entries ----
namespace {
__global__ void vertexCountKernel(const int *__restrict__ ids, long long n, int *__restrict__ count)
{
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e < n)
    atomicAdd(&count[ids[e]], 1);
}
__global__ void vertexFillKernel(const int *__restrict__ ids, long long n, const long long *__restrict__ off, int *__restrict__ cursor,
                                 long long *__restrict__ entries)
{
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e < n) {
    const int v                              = ids[e];
    entries[off[v] + atomicAdd(&cursor[v], 1)] = e;
  }
}
__global__ void vertexSortKernel(long long nVertices, const long long *__restrict__ off, long long *__restrict__ entries)
{
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= nVertices)
    return;
  const long long beg = off[v], end = off[v + 1];
  for (long long a = beg + 1; a < end; ++a) { // a handful of entries per vertex
    const long long x = entries[a];
    long long       b = a - 1;
    while (b >= beg && entries[b] > x) {
      entries[b + 1] = entries[b];
      --b;
    }
    entries[b + 1] = x;
  }
}
// entry e -> its cluster: the largest c with outOff[c] <= e
__device__ __forceinline__ long long clusterOfEntry(const long long *__restrict__ offs, long long nClusters, long long e)
{
  long long lo = 0, hi = nClusters - 1;
  while (lo < hi) {
    const long long mid = (lo + hi + 1) >> 1;
    if (offs[mid] <= e)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}
__global__ void detWeightSumKernel(PumSolveArgs a, long long nClusters, long long nVertices, const long long *__restrict__ off,
                                   const long long *__restrict__ entries, double *__restrict__ wsum)
{
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= nVertices)
    return;
  const double rinv = 1.0 / a.radius;
  double       s    = 0.0;
  for (long long k = off[v]; k < off[v + 1]; ++k) {
    const long long e = entries[k];
    const long long c = clusterOfEntry(a.outOff, nClusters, e);
    const double4   r = a.outRec[e];
    // computeWeight (SphericalVertexCluster.hpp:337-344): squared difference (centre, vertex)
    s += pumWeight(__dsqrt_rn(sqDist3(a.centers[3 * c], a.centers[3 * c + 1], a.centers[3 * c + 2], r.x, r.y, r.z)), rinv);
  }
  if (off[v + 1] > off[v])
    wsum[v] = s;
}
__global__ void blendGatherKernel(const double *__restrict__ stage, long long nVertices, const long long *__restrict__ off,
                                  const long long *__restrict__ entries, int dims, double *__restrict__ out)
{
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (v >= nVertices)
    return;
  for (int d = 0; d < dims; ++d) {
    double s = 0.0; // PartitionOfUnityMapping.hpp:352, 371: the output starts from zero, the clusters add in index order
    for (long long k = off[v]; k < off[v + 1]; ++k)
      s += stage[entries[k] * dims + d];
    out[v * dims + d] = s;
  }
}
} // namespace

void pumBuildVertexIndex(const int *ids, int64_t nEntries, int64_t nVertices, PumVertexIndex &ix, cudaStream_t s)
{
  ix.nVertices = nVertices;
  ix.off.alloc(nVertices + 1);
  ix.entries.alloc(std::max<int64_t>(nEntries, 1));
  DevBuf<int> count;
  count.alloc(std::max<int64_t>(nVertices, 1));
  RBF_CUDA(cudaMemsetAsync(count.p, 0, std::max<int64_t>(nVertices, 1) * sizeof(int), s));
  if (nEntries > 0) {
    vertexCountKernel<<<divUp(nEntries, 256), 256, 0, s>>>(ids, nEntries, count.p);
    RBF_LAUNCHED();
  }
  exclusiveScan(count.p, ix.off.p, nVertices, s);
  if (nEntries > 0) {
    RBF_CUDA(cudaMemsetAsync(count.p, 0, std::max<int64_t>(nVertices, 1) * sizeof(int), s));
    vertexFillKernel<<<divUp(nEntries, 256), 256, 0, s>>>(ids, nEntries, reinterpret_cast<const long long *>(ix.off.p), count.p,
                                                         reinterpret_cast<long long *>(ix.entries.p));
    RBF_LAUNCHED();
    vertexSortKernel<<<divUp(nVertices, 256), 256, 0, s>>>(nVertices, reinterpret_cast<const long long *>(ix.off.p), reinterpret_cast<long long *>(ix.entries.p));
    RBF_LAUNCHED();
  }
  RBF_CUDA(cudaGetLastError());
  RBF_CUDA(cudaStreamSynchronize(s)); // `count` is released on return
}

void pumDeterministicWeightSums(const PumSolveArgs &a, int64_t nClusters, const PumVertexIndex &ix, double *wsum, cudaStream_t s)
{
  if (ix.nVertices <= 0 || nClusters <= 0)
    return;
  detWeightSumKernel<<<divUp(ix.nVertices, 256), 256, 0, s>>>(a, nClusters, ix.nVertices, reinterpret_cast<const long long *>(ix.off.p),
                                                             reinterpret_cast<const long long *>(ix.entries.p), wsum);
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

void pumBlendGather(const double *stage, const PumVertexIndex &ix, int dims, double *out, cudaStream_t s)
{
  if (ix.nVertices <= 0)
    return;
  blendGatherKernel<<<divUp(ix.nVertices, 256), 256, 0, s>>>(stage, ix.nVertices, reinterpret_cast<const long long *>(ix.off.p),
                                                            reinterpret_cast<const long long *>(ix.entries.p), dims, out);
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

std::vector<long long> pumVertexHistogram(const double *dXyz, int64_t n, int axis, double lo, double binWidth, int nBins, cudaStream_t s)
{
  if (nBins > 4096)
    throw MapError(RBFMAP_ERR_INVALID, "too many histogram bins");
  DevBuf<int> hist;
  hist.alloc(nBins);
  RBF_CUDA(cudaMemsetAsync(hist.p, 0, nBins * sizeof(int), s));
  if (n > 0) {
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(smCount()) * 8)));
    vertexHistogramKernel<<<grid, 256, 0, s>>>(dXyz, n, axis, lo, binWidth > 0 ? 1.0 / binWidth : 0.0, nBins, hist.p);
    RBF_LAUNCHED();
  }
  std::vector<int> h(nBins);
  RBF_CUDA(cudaMemcpyAsync(h.data(), hist.p, nBins * sizeof(int), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  return std::vector<long long>(h.begin(), h.end());
}

int64_t pumSelectSlab(const double *dCenters, int64_t n, int axis, double lo, double hi, DevBuf<double> &out, DevBuf<int> &globalIdx, cudaStream_t s)
{
  if (n == 0) {
    out.alloc(3);
    globalIdx.alloc(1);
    return 0;
  }
  DevBuf<int> keep, pos;
  keep.alloc(n);
  pos.alloc(n + 1);
  const int grid = divUp(n, 256);
  slabFlagKernel<<<grid, 256, 0, s>>>(dCenters, n, axis, lo, hi, keep.p);
  RBF_LAUNCHED();
  exclusiveScan(keep.p, pos.p, n, s);
  int count = 0;
  RBF_CUDA(cudaMemcpyAsync(&count, pos.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  out.alloc(3 * static_cast<size_t>(std::max(count, 1)));
  globalIdx.alloc(std::max(count, 1));
  if (count > 0) {
    slabCompactKernel<<<grid, 256, 0, s>>>(n, keep.p, pos.p, dCenters, out.p, globalIdx.p);
    RBF_LAUNCHED();
  }
  RBF_CUDA(cudaGetLastError());
  return count;
}

// Cut positions (bin indices 0 = c_0 <= c_1 <= ... <= c_world = nBins) of a slab decomposition that minimises the largest
// number of centres a rank processes: rank r processes the bins [c_r - halo, c_{r+1} + halo).
void pumPartitionSlabs(const long long *hist, int nBins, int world, int haloBins, int *cuts)
{
  std::vector<long long> P(nBins + 1, 0);
  for (int b = 0; b < nBins; ++b)
    P[b + 1] = P[b] + hist[b];
  auto load = [&](int c0, int c1) { return P[std::min(nBins, c1 + haloBins)] - P[std::max(0, c0 - haloBins)]; };
  auto sweep = [&](long long T, int *out) { // greedy: every rank takes as many bins as T allows
    int c = 0;
    if (out)
      out[0] = 0;
    for (int r = 0; r < world; ++r) {
      int lo = c, hi = nBins; // largest e in [c, nBins] with load(c, e) <= T (load is monotone in e)
      if (load(c, c) > T)
        return false;
      while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        if (load(c, mid) <= T)
          lo = mid;
        else
          hi = mid - 1;
      }
      c = r == world - 1 ? (lo == nBins ? nBins : -1) : lo;
      if (c < 0)
        return false;
      if (out)
        out[r + 1] = c;
    }
    return true;
  };
  long long lo = 0, hi = P[nBins];
  while (lo < hi) {
    const long long mid = (lo + hi) / 2;
    if (sweep(mid, nullptr))
      hi = mid;
    else
      lo = mid + 1;
  }
  sweep(lo, cuts);
  // the greedy sweep fills the first ranks and may leave the last ones short; a backward pass would not change the
  // maximum, which is what bounds the step time
  cuts[world] = nBins;
}

void pumClassify(const PumMembership &m, int64_t nClusters, PumBuckets &b, rbfmap_stats &stats, cudaStream_t s)
{
  DevBuf<unsigned long long> acc;
  DevBuf<long long>          dStart;
  DevBuf<int>                cursor, sizeHist;
  acc.alloc(32);
  dStart.alloc(kNumSizeClasses + 1);
  cursor.alloc(kNumSizeClasses);
  sizeHist.alloc(kNumSizeClasses);
  RBF_CUDA(cudaMemsetAsync(acc.p, 0, 32 * sizeof(unsigned long long), s));
  RBF_CUDA(cudaMemsetAsync(cursor.p, 0, kNumSizeClasses * sizeof(int), s));
  RBF_CUDA(cudaMemsetAsync(sizeHist.p, 0, kNumSizeClasses * sizeof(int), s));
  classifyKernel<<<divUp(nClusters, 256), 256, 0, s>>>(m.nIn.p, m.nOut.p, nClusters, acc.p, sizeHist.p);
  RBF_LAUNCHED();
  unsigned long long h[32];
  int                hs[kNumSizeClasses];
  RBF_CUDA(cudaMemcpyAsync(h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaMemcpyAsync(hs, sizeHist.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  long long start[kNumSizeClasses + 1]; // per size class; the buckets are unions of consecutive size classes
  start[0] = 0;
  for (int k = 0; k < kNumSizeClasses; ++k) {
    start[k + 1]   = start[k] + hs[k];
    b.sizeStart[k] = start[k];
  }
  b.sizeStart[kNumSizeClasses] = start[kNumSizeClasses];
  int64_t run = 0;
  for (int k = 0; k < kNumBuckets; ++k) {
    b.start[k] = run;
    run += static_cast<int64_t>(h[k]);
    b.maxN[k] = static_cast<int>(h[19 + k]);
  }
  b.start[kNumBuckets] = run;
  stats.sum_n  = static_cast<int64_t>(h[13]);
  stats.sum_m  = static_cast<int64_t>(h[14]);
  stats.sum_n2 = static_cast<int64_t>(h[15]);
  stats.sum_n3 = static_cast<int64_t>(h[16]);
  stats.sum_mn = static_cast<int64_t>(h[17]);
  stats.max_n  = static_cast<int64_t>(h[18]);
  RBF_CUDA(cudaMemcpyAsync(dStart.p, start, sizeof(start), cudaMemcpyHostToDevice, s));
  b.order.alloc(std::max<int64_t>(nClusters, 1));
  b.entries.alloc(std::max<int64_t>(nClusters, 1));
  bucketScatterKernel<<<divUp(nClusters, 256), 256, 0, s>>>(m.nIn.p, m.nOut.p, reinterpret_cast<const long long *>(m.inOff.p),
                                                           reinterpret_cast<const long long *>(m.outOff.p), reinterpret_cast<const long long *>(m.matOff.p),
                                                           nClusters, dStart.p, cursor.p, b.order.p, b.entries.p);
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
  // no synchronisation: `start` is pageable memory (staged before cudaMemcpyAsync returns) and the temporaries are freed in
  // stream order
}
} // namespace rbfmap
