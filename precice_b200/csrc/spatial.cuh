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

  // bins `n` points (device AoS) with cell edge >= minCell; bounds are the mesh bounds
  void        build(const double *dXyz, int64_t n, const double bbMin[3], const double bbMax[3], double minCell, cudaStream_t s);
  BinGridView view() const
  {
    return BinGridView{origin[0], origin[1], origin[2], 1.0 / h[0], 1.0 / h[1], 1.0 / h[2], dims[0], dims[1], dims[2], cellStart.p, xyz.p, ids.p};
  }
  size_t bytes() const { return cellStart.bytes() + ids.bytes() + xyz.bytes(); }
};

// bounding box of n points (device AoS); returns min/max on the host (query/Index.cpp:369-391)
void computeBounds(const double *dXyz, int64_t n, double bbMin[3], double bbMax[3], cudaStream_t s);

// nearest vertex to each of nProbes host points: vertex ID with the smallest squared distance, ties -> lowest ID
// (query/Index.cpp:171-182; tie-break is implementation-defined in the reference)
void nearestVertices(const double *dXyz, int64_t n, const double *probes, int nProbes, int *nearestIds, cudaStream_t s);

// distance to the k-th nearest vertex (k >= 1, counting the query vertex itself) for each of nProbes
// vertices of the mesh, i.e. max over Index::getClosestVertices (query/Index.cpp:184-195) as consumed by
// impl/CreateClustering.hpp:259-268. `guess` is a starting search radius.
void kthNeighbourDistance(const double *dXyz, int64_t n, const int *probeIds, int nProbes, int k, double guess, double *result, cudaStream_t s);

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
