// pum.cuh — partition-of-unity mapping on the device: data structures shared by the clustering
// kernels (pum_cluster.cu), the per-cluster algebra kernels (pum_solve.cu) and the host driver (pum.cu).
#pragma once

#include <algorithm>
#include <functional>
#include <vector>

#include "basis.cuh"
#include "common.cuh"
#include "nccl_dyn.cuh"
#include "spatial.cuh"

namespace rbfmap {

// index offsets of the 8 / 26 lattice neighbours (impl/CreateClustering.hpp:54-82)
struct NeighbourOffsets {
  long long off[26];
  int       n;
};

// tentative centre lattice and its tags (impl/CreateClustering.hpp:411-496)
// the tentative centre lattice (impl/CreateClustering.hpp:438-463): per-axis coordinate arrays on the device
struct LatticeAxes {
  const double *a[3];
  int           n[3];
  double        start[3], invDist[3];
  // sharded mapping: only the layers rangeLo <= index[rangeAxis] <= rangeHi of the lattice are processed by this rank
  // (rangeAxis < 0: all); the centres outside count as empty
  int rangeAxis = -1, rangeLo = 0, rangeHi = -1;
};
__host__ __device__ inline bool latticeLayerInRange(const LatticeAxes &L, int x, int y, int z)
{
  if (L.rangeAxis < 0)
    return true;
  const int i = L.rangeAxis == 0 ? x : (L.rangeAxis == 1 ? y : z);
  return i >= L.rangeLo && i <= L.rangeHi;
}

struct PumClustering {
  int64_t        total  = 0; // lattice size
  double         radius = 0.0;
  LatticeAxes    axes{};
  DevBuf<int>    tag; // per lattice centre: != 0 = empty with respect to one of the meshes (tagEmptyLattice)
  DevBuf<int>    pos; // exclusive scan of (tag == 0): lattice index -> index in the live list
  // the live list: centres that passed the two emptiness tests (CreateClustering.hpp:468-471), ascending lattice index
  int64_t        nLive = 0;
  DevBuf<int>    liveIds; // lattice index
  DevBuf<double> centers; // nLive x 3
  DevBuf<int>    liveTag; // 1 = tagged for removal (duplicates, empty after the projection)

  void    initLattice(const LatticeAxes &axes, int tag0, cudaStream_t s);
  // tagEmptyClusters for centres on the lattice, from the vertex side: clears `bit` of tag[c] for every centre with a
  // vertex of `mesh` inside its sphere
  void    tagEmptyLattice(const BinGrid &mesh, int bit, cudaStream_t s);
  int64_t gatherLive(cudaStream_t s); // builds the live list, returns its length
  void    tagEmpty(const BinGrid &mesh, cudaStream_t s); // live list, arbitrary (projected) centres
  void    project(const BinGrid &inMesh, cudaStream_t s);
  void    tagDuplicates(const NeighbourOffsets &offs, double threshold, cudaStream_t s);
  int64_t compact(DevBuf<double> &out, cudaStream_t s); // returns the number of surviving centres
};

// CSR description of all clusters (impl/SphericalVertexCluster.hpp:153-162): sorted vertex IDs
struct PumMembership {
  DevBuf<int>     nIn, nOut;             // per cluster
  DevBuf<int64_t> tri;                   // n(n+1)/2 per cluster
  DevBuf<int64_t> inOff, outOff, matOff; // nClusters + 1 each
  DevBuf<int>     inIds, outIds;
  int64_t         sumIn = 0, sumOut = 0, sumTri = 0;
  size_t          bytes() const
  {
    return nIn.bytes() + nOut.bytes() + tri.bytes() + inOff.bytes() + outOff.bytes() + matOff.bytes() + inIds.bytes() + outIds.bytes();
  }
};

// Scratch of the single-pass membership (fixed-size slot of sorted vertex IDs per cluster and mesh). It belongs to the
// mapping object and survives clear(): re-allocating ~1 GB between the release and the re-allocation of the 9 GB factor
// buffer makes the stream-ordered pool split and re-create its large blocks every step (measured: +30 ms at 10M).
struct PumScratch {
  ScratchBuf<int> slotIn, slotOut;
  size_t      bytes() const { return slotIn.bytes() + slotOut.bytes(); }
};

// Slab of a sharded mapping: this rank owns what lies in lo <= x[axis] < hi
struct PumSlab {
  int    axis = 0;
  double lo = 0.0, hi = 0.0;
};

// outSlab (may be null): only the output vertices inside the slab are listed and weighted (sharded mapping, duplicate mode).
// weightsComplete (may be empty) runs after the weight sums of this rank's clusters are accumulated and before the
// normalised weights are derived from them (sharded mapping, exchange mode: all-reduce of the sums).
void    pumBuildMembership(const double *dCenters, int64_t nClusters, const BinGrid &inMesh, const BinGrid &outMesh, double radius, PumMembership &m,
                           PumScratch &scratch, double *wsum, int *wcount, const double *inXyz, const double *outXyz, DevBuf<double> &inRec,
                           DevBuf<double4> &outRec, cudaStream_t s, const PumSlab *outSlab = nullptr,
                           const std::function<void()> &weightsComplete = {}, bool fullMatrices = false);
// sharded mapping: histogram of the vertex coordinates along `axis`, selection of the centres inside [lo, hi) (order
// preserved; globalIdx = index in the list handed in), and the host-side choice of the cut positions
std::vector<long long> pumVertexHistogram(const double *dXyz, int64_t n, int axis, double lo, double binWidth, int nBins, cudaStream_t s);
int64_t pumSelectSlab(const double *dCenters, int64_t n, int axis, double lo, double hi, DevBuf<double> &out, DevBuf<int> &globalIdx, cudaStream_t s);
void    pumPartitionSlabs(const long long *hist, int nBins, int world, int haloBins, int *cuts /* world + 1 */);
bool    pumBuildPointMembership(const double *dCenters, int64_t nClusters, const BinGrid &points, double radius, PumScratch &scratch, DevBuf<int> &nPts,
                                DevBuf<int64_t> &off, DevBuf<int> &ids, double *wsum, int *wcount, cudaStream_t s);
// -1: every out-vertex (of the slab, if given) lies in a cluster; with a group the lowest index over all ranks
// device-only variant: atomicMin of the lowest unassigned vertex into *dFirst (0x7f7f7f7f = none); no synchronisation
void pumMarkFirstUnassigned(const double *wsum, const int *wcount, int64_t nOut, cudaStream_t s, const double *outXyz, const PumSlab *slab, int *dFirst);
// wsum / wcount: partition-of-unity weight sum of every out-vertex and the number of its covering clusters with weight zero
int64_t pumFirstUnassigned(const double *wsum, const int *wcount, int64_t nOut, cudaStream_t s, const double *outXyz = nullptr,
                           const PumSlab *slab = nullptr, const ShardGroup *group = nullptr);

// Per-cluster separated-polynomial record (RadialBasisFctSolver.hpp:377-410 condensed).
//  mode 0: no polynomial
//  mode 1: n >= p, full column rank. Basis [1, (x_a - c_a)/r ...] over the active axes `mask`,
//          G = inverse Gram matrix of that basis (row-major 4x4, leading p x p block used).
//  mode 2: n < p (under-determined, SphericalVertexCluster.hpp:170-172): G holds the p x n operator S with
//          beta = S f in the RAW basis [1, x_a ...] (S[q*4 + i]), i.e. ColPivHouseholderQR::solve.
struct PolyRec {
  double G[16];
  int    mode, mask, p, pad;
};

struct PumSolveArgs {
  RbfParams        rbf;
  int              dim;
  double           radius; // cluster radius
  const double    *inXyz, *outXyz; // solver-side meshes (device AoS)
  const double    *centers;
  const int       *nIn, *nOut;
  const long long *inOff, *outOff, *matOff;
  const int       *inIds, *outIds;
  double          *mat; // packed factors
  PolyRec         *poly;
  const double    *wsum;
  const int       *wcount;
  // just-in-time write mapping (completeJustInTimeMapping): accumulated A^T (w u) and V^T (w u) of a cache, laid out
  // [(inOff[c] + i) * dims + d] and [(4 c + k) * dims + d]; null for the ordinary conservative map
  const double *cacheAu = nullptr, *cachePoly = nullptr;
  // cluster-ordered out-vertex records (x, y, z, normalised partition-of-unity weight) at [outOff[c] + o]: one coalesced
  // 32-byte load per out-vertex in the map kernels instead of ID -> coordinates -> weight sum -> weight evaluation
  const double4 *outRec = nullptr;
  // cluster-ordered in-vertex coordinates at [(inOff[c] + i) * 3 + d] (factorisation and polynomial set-up read them
  // contiguously instead of ID -> coordinates)
  const double *inRec = nullptr;
  // execution-mode minimal-compute: stored evaluation matrices (pum_eval.cu), entry (lane vertex l, loop vertex k) of
  // cluster c at eval[evalOff[c] + k * ld + l]; null = re-evaluate the basis function in every map (minimal-memory)
  const double    *eval    = nullptr;
  const long long *evalOff = nullptr;
  // basis functions that are not strictly positive definite (pum_inverse.cu): `mat` holds the explicit inverse of every
  // cluster matrix, n x n doubles with S[j n + i] = (C^-1)(i, j), instead of the packed factor
  bool inverse = false;
  // deterministic blend (rbfmap_config.deterministic): the map kernels store the weighted result of CSR entry e (out-vertex
  // entries for consistent, in-vertex entries for conservative constraints) at blendStage[e * dims + d] instead of adding
  // it to the output with atomics; pumBlendGather then sums every vertex's entries in ascending cluster order
  double *blendStage = nullptr;
};
// doubles per cluster in `mat` (even count: 16-byte aligned for the TMA bulk copy)
__host__ __device__ inline long long pumMatSize(long long n, bool inverse) { return ((inverse ? n * n : n * (n + 1) / 2) + 1) & ~1LL; }

// Clusters are bucketed by size; every bucket is served by its own kernel instantiation:
//   buckets 0..7  : n <= 64,  8x8 threads,  NB = ceil(n/8)  = 1..8 register blocks per thread and dimension
//   buckets 8..11 : n <= 128, 16x16 threads, NB = ceil(n/16) = 5..8
//   bucket  12    : larger clusters, global-memory kernels
constexpr int    kNumBuckets = 13;
constexpr size_t kPumMaxSmem = 200 * 1024;
// Largest supported cluster (in-vertices): the tightest shared-memory bound of the global-memory kernels of bucket 12
// (conservative map, 3 components: 48 B x (n + 256) + 1 KB <= kPumMaxSmem). Checked at the end of computeMapping so that
// the caller does not learn about it in the first map() only; documented in include/rbfmap.h.
constexpr int kPumMaxClusterVertices = 3900;
// basis functions that are not strictly positive definite: the explicit inverse is built in shared memory (pum_inverse.cu)
constexpr int kPumMaxInverseVertices = 128;
__host__ __device__ inline int pumBucketOf(int n)
{
  if (n <= 64)
    return (max(n, 1) + 7) / 8 - 1;
  if (n <= 128)
    return 8 + (n + 15) / 16 - 5;
  return 12;
}
// What a per-cluster kernel needs to start, in one 48-byte record per bucket slot: without it a CTA walks
// list -> sizes / offsets -> vertex IDs -> coordinates, four dependent DRAM accesses of ~1.5 us each under load.
struct __align__(16) PumEntry {
  int       c, n, m, pad;
  long long inOff, outOff, matOff, pad2;
};

// Inside a bucket the clusters are sorted by their size n (size classes 0..128 and "more than 128"): the map kernels are
// launched per group of sizes, with exactly the shared memory those sizes need (the packed factor is 8 n(n+1)/2 bytes, so
// a launch sized for the largest cluster of a bucket of eight sizes would keep fewer clusters resident per SM).
constexpr int kNumSizeClasses = 130;
__host__ __device__ inline int pumSizeClassOf(int n) { return n > 128 ? 129 : (n < 0 ? 0 : n); }
struct PumBuckets {
  DevBuf<int>      order;             // cluster indices sorted by size class (hence by bucket)
  DevBuf<PumEntry> entries;           // the same order, with sizes and offsets
  int64_t     start[kNumBuckets + 1]; // offsets into `order`
  int         maxN[kNumBuckets];      // largest cluster of each bucket
  int64_t     sizeStart[kNumSizeClasses + 1]; // offsets into `order` per size class
};
// calls fn(firstEntry, count, maxN) for the size groups of bucket b (groups of `width` consecutive sizes)
template <typename F>
inline void pumForSizeGroups(const PumBuckets &bk, int b, int width, F &&fn)
{
  int lo, hi; // sizes of the bucket
  if (b < 8) {
    lo = b == 0 ? 0 : 8 * b + 1;
    hi = 8 * (b + 1);
  } else if (b < 12) {
    lo = 16 * (b - 4) + 1;
    hi = 16 * (b - 3);
  } else {
    lo = hi = 129;
  }
  for (int n0 = lo; n0 <= hi; n0 += width) {
    const int     n1    = std::min(hi, n0 + width - 1);
    const int64_t count = bk.sizeStart[n1 + 1] - bk.sizeStart[n0];
    if (count <= 0)
      continue;
    int maxN = n1;
    while (maxN > n0 && bk.sizeStart[maxN + 1] == bk.sizeStart[maxN])
      --maxN; // largest size class of the group that is populated
    fn(bk.sizeStart[n0], count, b == 12 ? bk.maxN[12] : maxN);
  }
}

// device-side classification + statistics (fills stats.sum_* and stats.max_n)
void pumClassify(const PumMembership &m, int64_t nClusters, PumBuckets &b, rbfmap_stats &stats, cudaStream_t s);

void pumPolySetup(const PumSolveArgs &a, int64_t nClusters, int polynomial, cudaStream_t s);
void pumBuildOutRecords(const PumSolveArgs &a, int64_t nClusters, double4 *rec, cudaStream_t s);
void pumBuildInRecords(const PumSolveArgs &a, int64_t sumIn, double *rec, cudaStream_t s);
// fail = device int (lowest failing cluster index or INT_MAX)
void pumFactor(const PumSolveArgs &a, const PumBuckets &b, int *fail, cudaStream_t s);
void pumInverse(const PumSolveArgs &a, const PumBuckets &b, int *fail, cudaStream_t s); // explicit inverses (pum_inverse.cu)
void pumFactorWarp(const PumSolveArgs &a, int bucket, const PumEntry *list, int64_t count, int *fail, cudaStream_t s);
// Deterministic blend: for every vertex the CSR entries that refer to it, ascending (= ascending
// cluster index = the order in which the reference's loop over the clusters adds to the output vector).
struct PumVertexIndex {
  DevBuf<int64_t> off;     // nVertices + 1
  DevBuf<int64_t> entries; // sum of cluster sizes
  int64_t         nVertices = 0;
};
void pumBuildVertexIndex(const int *ids, int64_t nEntries, int64_t nVertices, PumVertexIndex &ix, cudaStream_t s);
// weight sums of the out-vertices recomputed from the (sorted) entries, so that they do not depend on the order of atomics
void pumDeterministicWeightSums(const PumSolveArgs &a, int64_t nClusters, const PumVertexIndex &ix, double *wsum, cudaStream_t s);
void pumBlendGather(const double *stage, const PumVertexIndex &ix, int dims, double *out, cudaStream_t s);

// execution-mode minimal-compute (pum_eval.cu): sizes / offsets of the stored evaluation matrices, and their assembly
int64_t pumEvalOffsets(const PumMembership &m, int64_t nClusters, bool lanesAreOut, DevBuf<int64_t> &evalOff, cudaStream_t s);
void    pumBuildEvalMatrices(const PumSolveArgs &a, int64_t nClusters, bool lanesAreOut, const long long *evalOff, double *eval, cudaStream_t s);
void pumMapConsistent(const PumSolveArgs &a, const PumBuckets &b, const double *in, int dims, double *out, cudaStream_t s);
void pumMapConservative(const PumSolveArgs &a, const PumBuckets &b, const double *in, int dims, double *out, cudaStream_t s);
// just-in-time mapping (pum_jit.cu)
void pumJitUpdateCache(const PumSolveArgs &a, const PumBuckets &bk, const double *in, int dims, double *coef, double *poly, cudaStream_t s);
void pumJitInterpolateAt(const PumSolveArgs &a, const BinGridView &centers, const double *coords, int64_t nPts, int dims, const double *coef, const double *poly,
                         double *out, int *firstUnassigned, cudaStream_t s);
// cluster-centric evaluation for large point sets: per-cluster point lists (pumBuildPointMembership) + the map kernel's evaluation
void pumJitEvaluateClustered(const PumSolveArgs &a, int64_t nClusters, int maxN, const double *coords, const int *nPts, const long long *ptOff, const int *ptIds,
                             const double *wsum, const int *wcount, int dims, const double *coef, const double *poly, double *out, cudaStream_t s);
void pumJitAccumulateAt(const PumSolveArgs &a, const BinGridView &centers, const double *coords, int64_t nPts, int dims, const double *source, double *coef,
                        double *poly, int *firstUnassigned, cudaStream_t s);

} // namespace rbfmap
