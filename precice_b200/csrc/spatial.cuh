// spatial.cuh — uniform-grid neighbour search on the device.
//
// Replaces the host Boost.Geometry R*-tree (query/Index.cpp:171-253, 369-391) for exactly the
// queries the RBF hot path issues: bounds, nearest vertex, k nearest vertices, all vertices inside a
// sphere, any vertex inside a sphere. Sphere membership keeps the reference semantics: STRICT
// sqrt(d2) < radius on all three stored components (Index.cpp:235).
#pragma once

#include "common.cuh"

namespace rbfmap {

// POD view handed to kernels by value
struct BinGridView {
  double        ox, oy, oz;    // origin (min corner of the binned mesh)
  double        ihx, ihy, ihz; // inverse cell edge lengths
  int           nx, ny, nz;
  const int    *cellStart; // ncells + 1
  const double *xyz;       // binned coordinates, AoS n x 3, ascending vertex ID inside each cell
  const int    *ids;       // vertex ID of each binned point
};

struct BinGrid {
  double         origin[3] = {0, 0, 0}, h[3] = {1, 1, 1};
  int            dims[3] = {1, 1, 1};
  int64_t        n       = 0;
  DevBuf<int>    cellStart, ids;
  DevBuf<double> xyz;

  // bins `n` points (device AoS) with cell edge >= minCell; bounds are the mesh bounds.
  // regionAxis >= 0 (sharded mappings): only the points with regionLo <= x[regionAxis] <= regionHi are binned - the cell
  // geometry is that of the whole mesh, `n` becomes the number of binned points and the vertex IDs stay global.
  void        build(const double *dXyz, int64_t n, const double bbMin[3], const double bbMax[3], double minCell, cudaStream_t s, int regionAxis = -1,
                    double regionLo = 0.0, double regionHi = 0.0);
  BinGridView view() const
  {
    return BinGridView{origin[0], origin[1], origin[2], 1.0 / h[0], 1.0 / h[1], 1.0 / h[2], dims[0], dims[1], dims[2], cellStart.p, xyz.p, ids.p};
  }
  size_t bytes() const { return cellStart.bytes() + ids.bytes() + xyz.bytes(); }
};

// bounding box of n points (device AoS); returns min/max on the host (query/Index.cpp:369-391)
// device -> device copy of n doubles with a kernel (see spatial.cu)
void deviceCopy(double *dst, const double *src, int64_t n, cudaStream_t s);
void computeBounds(const double *dXyz, int64_t n, double bbMin[3], double bbMax[3], cudaStream_t s);

// nearest vertex to each of nProbes host points: vertex ID with the smallest squared distance, ties -> lowest ID
// (query/Index.cpp:171-182; tie-break is implementation-defined in the reference)
void nearestVertices(const double *dXyz, int64_t n, const double *probes, int nProbes, int *nearestIds, cudaStream_t s, double *nearestXyz = nullptr);

// distance to the k-th nearest vertex (k >= 1, counting the query vertex itself) for each of nProbes
// vertices of the mesh, i.e. max over Index::getClosestVertices (query/Index.cpp:184-195) as consumed by
// impl/CreateClustering.hpp:259-268. `guess` is a starting search radius.
void kthNeighbourDistance(const double *dXyz, int64_t n, const int *probeIds, int nProbes, int k, double guess, double *result, cudaStream_t s,
                          const double *probeXyz = nullptr);

#ifdef __CUDACC__
__device__ __forceinline__ int cellCoordUnclamped(double p, double o, double ih, int n)
{
  double t = floor((p - o) * ih);
  t        = fmin(fmax(t, -2.0), static_cast<double>(n) + 1.0);
  return static_cast<int>(t);
}

// Calls f(pos, valid) for every binned point in the (2H+1)^3 cell neighbourhood of (cx,cy,cz), warp-cooperatively:
// all 32 lanes call f together (valid == false on padding lanes) so f may use warp collectives.
// f returns true to stop early (must be warp-uniform).
template <int H = 1, typename F>
__device__ __forceinline__ void forNeighbourhoodWarp(const BinGridView &g, double cx, double cy, double cz, int lane, F &&f)
{
  const int ix = cellCoordUnclamped(cx, g.ox, g.ihx, g.nx);
  const int iy = cellCoordUnclamped(cy, g.oy, g.ihy, g.ny);
  const int iz = cellCoordUnclamped(cz, g.oz, g.ihz, g.nz);
  const int z0 = max(iz - H, 0), z1 = min(iz + H, g.nz - 1);
  if (z0 > z1)
    return;
  for (int x = max(ix - H, 0); x <= min(ix + H, g.nx - 1); ++x)
    for (int y = max(iy - H, 0); y <= min(iy + H, g.ny - 1); ++y) {
      const int row = (x * g.ny + y) * g.nz;
      const int beg = g.cellStart[row + z0], end = g.cellStart[row + z1 + 1];
      for (int base = beg; base < end; base += 32) {
        const int pos = base + lane;
        if (f(pos, pos < end))
          return;
      }
    }
}

// Same candidates, flattened: the (2H+1)^2 z-runs of the neighbourhood are looked up by (2H+1)^2 lanes at once (one
// memory latency instead of one per run), their lengths are prefix-summed across the warp, and the concatenated
// candidate list is walked 32 at a time, so every lane has work in every iteration (a run holds ~35 points at 50
// vertices per cluster: run-by-run the second iteration of each run is almost empty).
// f(pos, valid) is called by all 32 lanes; it returns true (warp-uniform) to stop early.
template <int H = 1, typename F>
__device__ __forceinline__ void forCandidatesWarp(const BinGridView &g, double cx, double cy, double cz, int lane, F &&f)
{
  constexpr int      W = 2 * H + 1, R = W * W;
  constexpr unsigned kFull = 0xffffffffu;
  static_assert(R <= 32, "one lane per z-run");
  const int ix = cellCoordUnclamped(cx, g.ox, g.ihx, g.nx);
  const int iy = cellCoordUnclamped(cy, g.oy, g.ihy, g.ny);
  const int iz = cellCoordUnclamped(cz, g.oz, g.ihz, g.nz);
  const int z0 = max(iz - H, 0), z1 = min(iz + H, g.nz - 1);
  int       beg = 0, len = 0;
  if (lane < R && z0 <= z1) {
    const int x = ix + lane / W - H, y = iy + lane % W - H;
    if (x >= 0 && x < g.nx && y >= 0 && y < g.ny) {
      const int row = (x * g.ny + y) * g.nz;
      beg           = g.cellStart[row + z0];
      len           = g.cellStart[row + z1 + 1] - beg;
    }
  }
  int incl = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o)
      incl += t;
  }
  const int total = __shfl_sync(kFull, incl, 31);
  const int excl  = incl - len;
  for (int base = 0; base < total; base += 32) {
    const int t = base + lane;
    int       r = 0; // number of runs that end at or before candidate t (binary search over the inclusive sums)
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int v = __shfl_sync(kFull, incl, r + step - 1);
      if (v <= t)
        r += step;
    }
    r             = min(r, 31);
    const int pos = __shfl_sync(kFull, beg, r) + (t - __shfl_sync(kFull, excl, r));
    if (f(pos, t < total))
      return;
  }
}

// single-thread variant: f(pos) for every binned point of the neighbourhood
template <typename F>
__device__ __forceinline__ void forNeighbourhoodThread(const BinGridView &g, double cx, double cy, double cz, F &&f)
{
  const int ix = cellCoordUnclamped(cx, g.ox, g.ihx, g.nx);
  const int iy = cellCoordUnclamped(cy, g.oy, g.ihy, g.ny);
  const int iz = cellCoordUnclamped(cz, g.oz, g.ihz, g.nz);
  const int z0 = max(iz - 1, 0), z1 = min(iz + 1, g.nz - 1);
  if (z0 > z1)
    return;
  for (int x = max(ix - 1, 0); x <= min(ix + 1, g.nx - 1); ++x)
    for (int y = max(iy - 1, 0); y <= min(iy + 1, g.ny - 1); ++y) {
      const int row = (x * g.ny + y) * g.nz;
      const int beg = g.cellStart[row + z0], end = g.cellStart[row + z1 + 1];
      for (int pos = beg; pos < end; ++pos)
        f(pos);
    }
}
#endif

} // namespace rbfmap
