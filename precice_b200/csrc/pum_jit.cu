// pum_jit.cu — just-in-time mapping for rbf-pum-direct on the device (sm_100a).
//
// Replaces PartitionOfUnityMapping::updateMappingDataCache / mapConsistentAt / mapConservativeAt
// (PartitionOfUnityMapping.hpp:382-476) with SphericalVertexCluster::computeCacheData / interpolateAt /
// addWriteDataToCache (impl/SphericalVertexCluster.hpp:255-292) and RadialBasisFctSolver::computeCacheData /
// interpolateAt / addWriteDataToCache (RadialBasisFctSolver.hpp:470-560). The reference forbids just-in-time mapping
// on its GPU path (PartitionOfUnityMapping.hpp:211); SURVEY.md §8(f) rank 2.
//
// Cache layout (one per data, owned by the caller's rbfmap_cache): coef[(inOff[c] + i) * dims + d] = coefficient of
// in-vertex i of cluster c (consistent: lambda; conservative: accumulated A^T (w u)), poly[(4 c + k) * dims + d] =
// polynomial coefficient k of cluster c in the cluster-centred basis of PolyRec (consistent: beta; conservative:
// accumulated V^T (w u)).
#include "pum_map.cuh"

namespace rbfmap {
namespace {

constexpr int kMaxCover = 64; // clusters that may cover one point (typically 4-8)

// ---- cache update: gather, polynomial fit, lambda = C^-1 (f - Q beta); one warp per cluster ----
template <int NC, int RPL>
__global__ void __launch_bounds__(32) pumCoefficientKernel(PumSolveArgs args, const int *__restrict__ list, const double *__restrict__ in, int dims, int c0,
                                                           double *__restrict__ coef, double *__restrict__ polyOut)
{
  constexpr int LDV = TileLd<NC>::value;
  constexpr int NT  = 32;
  extern __shared__ __align__(16) double dyn[];
  __shared__ __align__(8) unsigned long long mbar;

  const int       c   = list[blockIdx.x];
  const int       n   = args.nIn[c];
  const int       tid = threadIdx.x;
  const long long off = args.inOff[c];
  const int      *ids = args.inIds + off;
  const double    cx = args.centers[3 * c], cy = args.centers[3 * c + 1], cz = args.centers[3 * c + 2];
  const double    rinv   = 1.0 / args.radius;
  const int       triPad = static_cast<int>(paddedTri(n));
  double *ms  = dyn;
  double *vt  = dyn + triPad;
  double *red = vt + n * LDV;

  if (tid == 0)
    mbarInit(&mbar, 1);
  __syncwarp();
  if (tid == 0)
    bulkLoad(ms, args.mat + args.matOff[c], static_cast<unsigned>(triPad) * 8u, &mbar);

  for (int i = tid; i < n; i += NT) { // SphericalVertexCluster.hpp:260-265
    const long long id = ids[i];
    const double   *p  = args.inXyz + 3 * id;
    vt[i * LDV]        = p[0];
    vt[i * LDV + 1]    = p[1];
    vt[i * LDV + 2]    = p[2];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      vt[i * LDV + 3 + cc] = in[id * dims + c0 + cc];
  }
  __syncwarp();

  // RadialBasisFctSolver.hpp:474-478
  const int polyMode = args.poly[c].mode, polyMask = args.poly[c].mask, polyRefine = args.poly[c].pad;
  double    beta[4][NC];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      beta[q][cc] = 0.0;
  if (polyMode == 1) {
    const double *G = args.poly[c].G;
#pragma unroll 1
    for (int pass = 0; pass <= polyRefine; ++pass) {
      double t[4 * NC];
#pragma unroll
      for (int e = 0; e < 4 * NC; ++e)
        t[e] = 0.0;
      for (int i = tid; i < n; i += NT) {
        double q[4];
        polyRow(1, polyMask, vt[i * LDV], vt[i * LDV + 1], vt[i * LDV + 2], cx, cy, cz, rinv, q);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double r = vt[i * LDV + 3 + cc];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            r -= q[k] * beta[k][cc];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            t[k * NC + cc] += q[k] * r;
        }
      }
      blockReduceSum<4 * NC>(t, red, NT);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < 4; ++l)
            s += G[4 * k + l] * t[l * NC + cc];
          beta[k][cc] += s;
        }
    }
  } else if (polyMode == 2) {
    const double *G = args.poly[c].G;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        double s = 0.0;
        for (int i = 0; i < n && i < 3; ++i)
          s += G[4 * k + i] * vt[i * LDV + 3 + cc];
        beta[k][cc] = s;
      }
  }
  if (polyMode != 0) {
    __syncwarp();
    for (int i = tid; i < n; i += NT) {
      double q[4];
      polyRow(polyMode, polyMask, vt[i * LDV], vt[i * LDV + 1], vt[i * LDV + 2], cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        double r = vt[i * LDV + 3 + cc];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          r -= q[k] * beta[k][cc];
        vt[i * LDV + 3 + cc] = r;
      }
    }
    __syncwarp();
  }
  if (tid < 4) {
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) {
      double b = beta[0][cc];
      if (tid == 1) b = beta[1][cc];
      if (tid == 2) b = beta[2][cc];
      if (tid == 3) b = beta[3][cc];
      polyOut[(static_cast<long long>(c) * 4 + tid) * dims + c0 + cc] = b;
    }
  }
  // RadialBasisFctSolver.hpp:481
  mbarWait(&mbar, 0);
  warpSolve<NC, LDV, RPL>(ms, n, vt + 3, tid);
  __syncwarp();
  for (int i = tid; i < n; i += NT) {
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      coef[(off + i) * dims + c0 + cc] = vt[i * LDV + 3 + cc];
  }
}

// ---- clusters that cover a point and their normalised weights (PartitionOfUnityMapping.hpp:290-341); warp-cooperative ----
// returns the number of covering clusters (0: unassigned); ids / w live in the warp's shared memory
__device__ __forceinline__ int coveringClusters(const BinGridView &gc, double x, double y, double z, double radius, int *ids, double *w, int lane)
{
  const double rinv = 1.0 / radius;
  int          cnt  = 0;
  forCandidatesWarp<2>(gc, x, y, z, lane, [&](int pos, bool valid) {
    double d = 1.7976931348623157e308;
    if (valid) // computeWeight (SphericalVertexCluster.hpp:337-344): squared difference (centre, vertex)
      d = __dsqrt_rn(sqDist3(gc.xyz[3 * pos], gc.xyz[3 * pos + 1], gc.xyz[3 * pos + 2], x, y, z));
    const bool     hit = d < radius;
    const unsigned m   = __ballot_sync(0xffffffffu, hit);
    const int      at  = cnt + __popc(m & ((1u << lane) - 1u));
    if (hit && at < kMaxCover) {
      ids[at] = gc.ids[pos];
      w[at]   = pumWeight(d, rinv);
    }
    cnt += __popc(m);
    return false;
  });
  __syncwarp();
  cnt = min(cnt, kMaxCover);
  // the reference sums the weights in the order of its R-tree query; here: ascending cluster index (deterministic)
  if (lane == 0) {
    for (int a = 1; a < cnt; ++a) { // insertion sort, cnt is small
      const int    ia = ids[a];
      const double wa = w[a];
      int          b  = a - 1;
      while (b >= 0 && ids[b] > ia) {
        ids[b + 1] = ids[b];
        w[b + 1]   = w[b];
        --b;
      }
      ids[b + 1] = ia;
      w[b + 1]   = wa;
    }
    double sum = 0.0;
    for (int a = 0; a < cnt; ++a)
      sum = __dadd_rn(sum, w[a]);
    if (sum <= 0.0) { // :329-333
      for (int a = 0; a < cnt; ++a)
        w[a] = __ddiv_rn(1.0, static_cast<double>(cnt));
    } else {
      for (int a = 0; a < cnt; ++a)
        w[a] = __ddiv_rn(w[a], sum);
    }
  }
  __syncwarp();
  return cnt;
}

constexpr int kJitWarps = 4;

// ---- mapConsistentAt: one warp per point; values(v) = sum_c w_c(v) [ sum_j phi(|v - x_j|) lambda_cj + poly_c(v) ] ----
template <int KIND>
__global__ void __launch_bounds__(32 * kJitWarps) pumInterpolateAtKernel(PumSolveArgs args, BinGridView gc, const double *__restrict__ coords, long long nPts,
                                                                         int dims, const double *__restrict__ coef, const double *__restrict__ poly,
                                                                         double *__restrict__ out, int *__restrict__ firstUnassigned)
{
  __shared__ int    sIds[kJitWarps][kMaxCover];
  __shared__ double sW[kJitWarps][kMaxCover];
  const int       lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const long long v    = static_cast<long long>(blockIdx.x) * kJitWarps + wp;
  if (v >= nPts)
    return;
  const double    x = coords[3 * v], y = coords[3 * v + 1], z = coords[3 * v + 2];
  const RbfParams rbf  = args.rbf;
  const double    rinv = 1.0 / args.radius;
  const int       cnt  = coveringClusters(gc, x, y, z, args.radius, sIds[wp], sW[wp], lane);
  if (cnt == 0) {
    if (lane == 0)
      atomicMin(firstUnassigned, static_cast<int>(v));
    return;
  }
  for (int d0 = 0; d0 < dims; d0 += 3) {
    const int nd = min(3, dims - d0);
    double    total[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < cnt; ++a) {
      const int       c   = sIds[wp][a];
      const int       n   = args.nIn[c];
      const long long off = args.inOff[c];
      const int      *ids = args.inIds + off;
      double          acc[3] = {0.0, 0.0, 0.0};
      for (int i = lane; i < n; i += 32) { // RadialBasisFctSolver.hpp:497-504 (all three axes in the distance)
        const double *p   = args.inXyz + 3 * static_cast<long long>(ids[i]);
        const double  phi = evalBasisK<KIND>(distSqrt(sqDistD<true>(x, y, z, p[0], p[1], p[2])), rbf);
        for (int d = 0; d < nd; ++d)
          acc[d] = fma(phi, coef[(off + i) * dims + d0 + d], acc[d]);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
      const int polyMode = args.poly[c].mode;
      if (polyMode != 0) { // :508-518 in the basis of the cluster's polynomial record
        double q[4];
        polyRow(polyMode, args.poly[c].mask, x, y, z, args.centers[3 * c], args.centers[3 * c + 1], args.centers[3 * c + 2], rinv, q);
        for (int d = 0; d < nd; ++d)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            acc[d] += q[k] * poly[(static_cast<long long>(c) * 4 + k) * dims + d0 + d];
      }
      for (int d = 0; d < nd; ++d)
        total[d] += sW[wp][a] * acc[d]; // PartitionOfUnityMapping.hpp:471-472
    }
    if (lane == 0)
      for (int d = 0; d < nd; ++d)
        out[v * dims + d0 + d] = total[d];
  }
}

// ---- mapConservativeAt: one warp per point; Au_c += phi(|v - x_j|) w_c u, eps_c += V(v)^T w_c u ----
template <int KIND>
__global__ void __launch_bounds__(32 * kJitWarps) pumAccumulateAtKernel(PumSolveArgs args, BinGridView gc, const double *__restrict__ coords, long long nPts,
                                                                        int dims, const double *__restrict__ source, double *__restrict__ coef,
                                                                        double *__restrict__ poly, int *__restrict__ firstUnassigned)
{
  __shared__ int    sIds[kJitWarps][kMaxCover];
  __shared__ double sW[kJitWarps][kMaxCover];
  const int       lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const long long v    = static_cast<long long>(blockIdx.x) * kJitWarps + wp;
  if (v >= nPts)
    return;
  const double    x = coords[3 * v], y = coords[3 * v + 1], z = coords[3 * v + 2];
  const RbfParams rbf  = args.rbf;
  const double    rinv = 1.0 / args.radius;
  const int       cnt  = coveringClusters(gc, x, y, z, args.radius, sIds[wp], sW[wp], lane);
  if (cnt == 0) {
    if (lane == 0)
      atomicMin(firstUnassigned, static_cast<int>(v));
    return;
  }
  for (int a = 0; a < cnt; ++a) {
    const int       c   = sIds[wp][a];
    const int       n   = args.nIn[c];
    const long long off = args.inOff[c];
    const int      *ids = args.inIds + off;
    const double    w   = sW[wp][a];
    for (int i = lane; i < n; i += 32) { // RadialBasisFctSolver.hpp:536-543
      const double *p   = args.inXyz + 3 * static_cast<long long>(ids[i]);
      const double  phi = evalBasisK<KIND>(distSqrt(sqDistD<true>(x, y, z, p[0], p[1], p[2])), rbf);
      for (int d = 0; d < dims; ++d)
        atomicAdd(&coef[(off + i) * dims + d], phi * (w * source[v * dims + d]));
    }
    const int polyMode = args.poly[c].mode;
    if (polyMode != 0 && lane < 4) { // :547-559 in the basis of the cluster's polynomial record
      double q[4];
      polyRow(polyMode, args.poly[c].mask, x, y, z, args.centers[3 * c], args.centers[3 * c + 1], args.centers[3 * c + 2], rinv, q);
      double qk = q[0];
      if (lane == 1) qk = q[1];
      if (lane == 2) qk = q[2];
      if (lane == 3) qk = q[3];
      for (int d = 0; d < dims; ++d)
        atomicAdd(&poly[(static_cast<long long>(c) * 4 + lane) * dims + d], qk * (w * source[v * dims + d]));
    }
  }
}

// ---- cluster-centric mapConsistentAt for large point sets: one warp per cluster evaluates all query points that lie in
// it (the evaluation loop of pumMapConsistentKernel with the cached coefficients instead of a solve) ----
template <int KIND, int NC>
__global__ void __launch_bounds__(32, 16) pumEvaluateClusteredKernel(PumSolveArgs args, const double *__restrict__ coords, const int *__restrict__ nPts,
                                                                 const long long *__restrict__ ptOff, const int *__restrict__ ptIds,
                                                                 const double *__restrict__ wsum, const int *__restrict__ wcount, int dims, int c0,
                                                                 const double *__restrict__ coef, const double *__restrict__ poly, double *__restrict__ out)
{
  constexpr int LDV = TileLd<NC>::value;
  extern __shared__ __align__(16) double vt[];
  const int c = blockIdx.x;
  const int m = nPts[c];
  if (m == 0)
    return;
  const int       n    = args.nIn[c];
  const int       tid  = threadIdx.x;
  const long long off  = args.inOff[c];
  const int      *oids = ptIds + ptOff[c];
  const RbfParams rbf  = args.rbf;
  const double    cx = args.centers[3 * c], cy = args.centers[3 * c + 1], cz = args.centers[3 * c + 2];
  const double    rinv  = 1.0 / args.radius;
  const double    invR2 = rbf.p1 * rbf.p1;
  const int       polyMode = args.poly[c].mode, polyMask = args.poly[c].mask;
  for (int i = tid; i < n; i += 32) {
    const double *p = args.inRec + 3 * (off + i); // cluster-ordered coordinates
    vt[i * LDV]     = p[0];
    vt[i * LDV + 1] = p[1];
    vt[i * LDV + 2] = p[2];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      vt[i * LDV + 3 + cc] = coef[(off + i) * dims + c0 + cc];
  }
  double beta[4][NC];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      beta[k][cc] = polyMode != 0 ? poly[(static_cast<long long>(c) * 4 + k) * dims + c0 + cc] : 0.0;
  __syncwarp();
  for (int o0 = 0; o0 < m; o0 += 32) {
    const int rem   = m - o0;
    const int split = rem <= 16 ? (rem <= 4 ? 8 : rem <= 8 ? 4 : 2) : 1;
    const int per   = 32 / split;
    const int ol = tid % per, sub = tid / per;
    const bool have = ol < rem;
    long long  gid = 0;
    double     x = 0.0, y = 0.0, z = 0.0;
    if (have) {
      gid             = oids[o0 + ol];
      const double *p = coords + 3 * gid;
      x = p[0], y = p[1], z = p[2];
    }
    double acc[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      acc[cc] = 0.0;
    if (have) {
      const double *vp = vt + sub * LDV;
      const int     stride = split * LDV;
      int j = sub;
      for (; j + 3 * split < n; j += 4 * split, vp += 4 * stride) {
        double bx[4], by[4], bz[4], cf[4][NC], phi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double2 xy = *reinterpret_cast<const double2 *>(vp + u * stride);
          const double2 zb = *reinterpret_cast<const double2 *>(vp + u * stride + 2);
          bx[u] = xy.x, by[u] = xy.y, bz[u] = zb.x, cf[u][0] = zb.y;
          if (NC > 1) {
            const double2 b12 = *reinterpret_cast<const double2 *>(vp + u * stride + 4);
            cf[u][1]          = b12.x;
            if (NC > 2)
              cf[u][2] = b12.y;
          }
        }
        pairBasisBatch<KIND, true, 4>(x, y, z, bx, by, bz, rbf, invR2, phi);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          acc[cc] += fma(phi[0], cf[0][cc], phi[1] * cf[1][cc]) + fma(phi[2], cf[2][cc], phi[3] * cf[3][cc]);
      }
      for (; j < n; j += split, vp += stride) {
        const double2 xy  = *reinterpret_cast<const double2 *>(vp);
        const double2 zb  = *reinterpret_cast<const double2 *>(vp + 2);
        const double  phi = pairBasis<KIND, true>(x, y, z, xy.x, xy.y, zb.x, rbf, invR2);
        acc[0]            = fma(phi, zb.y, acc[0]);
        if (NC > 1) {
          const double2 b12 = *reinterpret_cast<const double2 *>(vp + 4);
          acc[1]            = fma(phi, b12.x, acc[1]);
          if (NC > 2)
            acc[2] = fma(phi, b12.y, acc[2]);
        }
      }
    }
    if (split > 1) {
      for (int o = per; o < 32; o <<= 1) {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          acc[cc] += __shfl_xor_sync(0xffffffffu, acc[cc], o);
      }
    }
    if (have && sub == 0) {
      if (polyMode != 0) {
        double q[4];
        polyRow(polyMode, polyMask, x, y, z, cx, cy, cz, rinv, q);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            acc[cc] += q[k] * beta[k][cc];
      }
      const double w = normalizedWeight(x, y, z, cx, cy, cz, rinv, wsum[gid], wcount + gid);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        atomicAdd(&out[gid * dims + c0 + cc], acc[cc] * w);
    }
  }
}

template <int KIND>
void evaluateClustered(const PumSolveArgs &a, int64_t nClusters, int maxN, const double *coords, const int *nPts, const long long *ptOff, const int *ptIds,
                       const double *wsum, const int *wcount, int dims, const double *coef, const double *poly, double *out, cudaStream_t s)
{
  for (int c0 = 0; c0 < dims;) {
    const int nc = std::min(3, dims - c0);
    const unsigned grid = static_cast<unsigned>(nClusters);
    if (nc == 1) {
      const size_t sm = static_cast<size_t>(TileLd<1>::value) * maxN * sizeof(double);
      pumEvaluateClusteredKernel<KIND, 1><<<grid, 32, sm, s>>>(a, coords, nPts, ptOff, ptIds, wsum, wcount, dims, c0, coef, poly, out);
    } else if (nc == 2) {
      const size_t sm = static_cast<size_t>(TileLd<2>::value) * maxN * sizeof(double);
      pumEvaluateClusteredKernel<KIND, 2><<<grid, 32, sm, s>>>(a, coords, nPts, ptOff, ptIds, wsum, wcount, dims, c0, coef, poly, out);
    } else {
      const size_t sm = static_cast<size_t>(TileLd<3>::value) * maxN * sizeof(double);
      pumEvaluateClusteredKernel<KIND, 3><<<grid, 32, sm, s>>>(a, coords, nPts, ptOff, ptIds, wsum, wcount, dims, c0, coef, poly, out);
    }
    RBF_LAUNCHED();
    c0 += nc;
  }
}

template <int NC>
void launchCoefficient(const PumSolveArgs &a, int bucket, const int *list, int64_t count, int maxN, const double *in, int dims, int c0, double *coef, double *poly,
                       cudaStream_t s)
{
  const size_t sm = (static_cast<size_t>(paddedTri(maxN)) + static_cast<size_t>(TileLd<NC>::value) * maxN + 16) * sizeof(double);
  if (bucket >= 12 || sm > kPumMaxSmem)
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct just-in-time mapping: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size (128).");
  if (bucket < 8) {
    allowSmem(pumCoefficientKernel<NC, 2>, sm);
    pumCoefficientKernel<NC, 2><<<static_cast<unsigned>(count), 32, sm, s>>>(a, list, in, dims, c0, coef, poly);
  } else {
    allowSmem(pumCoefficientKernel<NC, 4>, sm);
    pumCoefficientKernel<NC, 4><<<static_cast<unsigned>(count), 32, sm, s>>>(a, list, in, dims, c0, coef, poly);
  }
  RBF_LAUNCHED();
}

} // namespace

void pumJitUpdateCache(const PumSolveArgs &a, const PumBuckets &bk, const double *in, int dims, double *coef, double *poly, cudaStream_t s)
{
  for (int b = 0; b < kNumBuckets; ++b) {
    const int64_t count = bk.start[b + 1] - bk.start[b];
    if (count <= 0)
      continue;
    const int *list = bk.order.p + bk.start[b];
    for (int c0 = 0; c0 < dims;) {
      const int nc = std::min(3, dims - c0);
      if (nc == 1)
        launchCoefficient<1>(a, b, list, count, bk.maxN[b], in, dims, c0, coef, poly, s);
      else if (nc == 2)
        launchCoefficient<2>(a, b, list, count, bk.maxN[b], in, dims, c0, coef, poly, s);
      else
        launchCoefficient<3>(a, b, list, count, bk.maxN[b], in, dims, c0, coef, poly, s);
      c0 += nc;
    }
  }
  RBF_CUDA(cudaGetLastError());
}

void pumJitInterpolateAt(const PumSolveArgs &a, const BinGridView &centers, const double *coords, int64_t nPts, int dims, const double *coef, const double *poly,
                         double *out, int *firstUnassigned, cudaStream_t s)
{
  if (nPts <= 0)
    return;
  const unsigned grid = static_cast<unsigned>((nPts + kJitWarps - 1) / kJitWarps);
  RBF_DISPATCH_BASIS(a.rbf.kind, (pumInterpolateAtKernel<KIND><<<grid, 32 * kJitWarps, 0, s>>>(a, centers, coords, nPts, dims, coef, poly, out, firstUnassigned)));
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

void pumJitEvaluateClustered(const PumSolveArgs &a, int64_t nClusters, int maxN, const double *coords, const int *nPts, const long long *ptOff, const int *ptIds,
                             const double *wsum, const int *wcount, int dims, const double *coef, const double *poly, double *out, cudaStream_t s)
{
  if (nClusters <= 0)
    return;
  RBF_DISPATCH_HOT_BASIS(a.rbf.kind, evaluateClustered<KIND>(a, nClusters, maxN, coords, nPts, ptOff, ptIds, wsum, wcount, dims, coef, poly, out, s));
  RBF_CUDA(cudaGetLastError());
}

void pumJitAccumulateAt(const PumSolveArgs &a, const BinGridView &centers, const double *coords, int64_t nPts, int dims, const double *source, double *coef,
                        double *poly, int *firstUnassigned, cudaStream_t s)
{
  if (nPts <= 0)
    return;
  const unsigned grid = static_cast<unsigned>((nPts + kJitWarps - 1) / kJitWarps);
  RBF_DISPATCH_BASIS(a.rbf.kind, (pumAccumulateAtKernel<KIND><<<grid, 32 * kJitWarps, 0, s>>>(a, centers, coords, nPts, dims, source, coef, poly, firstUnassigned)));
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

} // namespace rbfmap
