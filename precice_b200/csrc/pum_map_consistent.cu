// pum_map_consistent.cu — per-cluster consistent map: gather, separated polynomial, solve, on-the-fly evaluation and
// partition-of-unity blending in one kernel (sm_100a).
//
// Replaces, per cluster, SphericalVertexCluster::mapConsistent (impl/SphericalVertexCluster.hpp:296-335) with
// RadialBasisFctSolver::solveConsistent (RadialBasisFctSolver.hpp:441-468) and the evaluation matrix of buildMatrixA
// (:209-242), which is never stored: A(i,j) = phi(|out_i - in_j|) is re-evaluated from the vertex tile in shared memory.
// (Kokkos prior art: do_batched_solve, device/KokkosPUMKernels_Impl.hpp:491-814.)
#include "pum_map.cuh"

#include <cstdio>
#include <cstdlib>

namespace rbfmap {
#ifdef RBF_PHASE_TIMING
__device__ unsigned long long g_mapPhase[8]; // cycles: 0 start->gathered, 1 poly, 2 solve, 3 eval, 4 clusters
#define RBF_STAMP(k)                                                                   \
  do {                                                                                 \
    const long long now_ = clock64();                                                  \
    if (threadIdx.x == 0)                                                              \
      atomicAdd(&g_mapPhase[k], static_cast<unsigned long long>(now_ - stamp_));       \
    stamp_ = now_;                                                                     \
  } while (0)
#else
#define RBF_STAMP(k)
#endif
namespace {

// STORED: execution-mode minimal-compute, the evaluation matrix is read from args.eval (pum_eval.cu) instead of being
// re-evaluated (KIND and DIM3 are then irrelevant: instantiated with KIND_RUNTIME / true only)
template <int KIND, int NC, int NT, bool STAGE_M, bool DIM3, bool STORED = false>
__global__ void __launch_bounds__(NT, NT == 32 ? 14 : 1) pumMapConsistentKernel(PumSolveArgs args, const PumEntry *__restrict__ list, const double *__restrict__ in, int dims, int c0,
                                                             double *__restrict__ out)
{
  constexpr int LDV = TileLd<NC>::value;
  extern __shared__ __align__(16) double dyn[];
  __shared__ __align__(8) unsigned long long mbar;

#ifdef RBF_PHASE_TIMING
  long long stamp_ = clock64();
#endif
  const PumEntry  en   = list[blockIdx.x];
  const int       c    = en.c;
  const int       n    = en.n;
  const int       m    = en.m;
  const int       tid  = threadIdx.x;
  if (m == 0)
    return; // nothing to evaluate (uniform per CTA)
  const int      *ids  = args.inIds + en.inOff;
  const int      *oids = args.outIds + en.outOff;
  const double   *mg   = args.mat + en.matOff;
  const RbfParams rbf  = args.rbf;
  const double    cx = args.centers[3 * c], cy = args.centers[3 * c + 1], cz = args.centers[3 * c + 2];
  const double    rinv = 1.0 / args.radius;
  const double    invR2 = rbf.p1 * rbf.p1;
  const int       triPad = STAGE_M ? static_cast<int>(pumMatSize(n, args.inverse)) : 0;

  double *ms  = dyn;
  double *vt  = dyn + triPad;
  double *red = vt + n * LDV;

  // ---- stage the packed factor with one TMA bulk copy; it lands while the tile is gathered and the polynomial fitted ----
  if (STAGE_M) {
    if (tid == 0)
      mbarInit(&mbar, 1);
    __syncthreads();
    if (tid == 0)
      bulkLoad(ms, mg, static_cast<unsigned>(triPad) * 8u, &mbar);
  }
  // stored evaluation matrix: on its way into L2 while the coefficients are computed (the evaluation loop would
  // otherwise wait for DRAM once per group of columns)
  if (STORED && tid == 0)
    bulkPrefetchL2(args.eval + args.evalOff[c], static_cast<unsigned>(((m + 1) & ~1) * n) * 8u);

  // ---- gather (SphericalVertexCluster.hpp:312-320) ----
  for (int i = tid; i < n; i += NT) {
    const long long id = ids[i];
    const double   *p  = args.inRec + 3 * (en.inOff + i); // cluster-ordered copy: contiguous
    vt[i * LDV]        = p[0];
    vt[i * LDV + 1]    = p[1];
    vt[i * LDV + 2]    = p[2];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      vt[i * LDV + 3 + cc] = in[id * dims + c0 + cc];
  }
  __syncthreads();

  RBF_STAMP(0);
  // ---- separated polynomial: beta = lstsq(Q, f); f <- f - Q beta (RadialBasisFctSolver.hpp:446-449) ----
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
    for (int pass = 0; pass <= polyRefine; ++pass) { // normal equations (+ one refinement step on the residual)
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
    __syncthreads();
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
    __syncthreads();
  }

  RBF_STAMP(1);
  // ---- lambda = C^-1 f ----
  if (args.inverse) { // basis function not strictly positive definite: explicit inverse (pum_inverse.cu), uniform branch
    if (STAGE_M)
      mbarWait(&mbar, 0);
    inverseApply<NC, LDV>(STAGE_M ? ms : mg, n, vt + 3, tid, NT);
  } else if (STAGE_M) {
    if (tid < 32) { // one warp solves; the others wait at the barrier
      mbarWait(&mbar, 0);
      warpSolve<NC, LDV, (NT == 32 ? 2 : NT / 32)>(ms, n, vt + 3, tid);
    }
    __syncthreads();
  } else {
    blockedSolve<NC, LDV, long long>(mg, n, vt + 3, tid, NT);
  }

  RBF_STAMP(2);
  // ---- out = A lambda + V beta, weighted accumulation (SphericalVertexCluster.hpp:322-334) ----
  // One out-vertex per lane and pass. When the last pass has 16 or fewer vertices left, 2 / 4 / 8 lanes share a vertex
  // (each takes every 2nd / 4th / 8th in-vertex, partial sums are combined with shuffles) so that no lane idles.
  for (int o0 = 0; o0 < m; o0 += NT) {
    const int rem   = m - o0;
    const int split = (!STORED && NT == 32 && rem <= 16) ? (rem <= 4 ? 8 : rem <= 8 ? 4 : 2) : 1;
    const int per   = NT / split;
    const int ol = tid % per, sub = tid / per;
    const bool have = ol < rem;
    long long  gid = 0;
    double     x = 0.0, y = 0.0, z = 0.0, wNorm = 0.0;
    if (have) {
      const double4 r = args.outRec[en.outOff + o0 + ol]; // x, y, z, normalised weight: one coalesced load
      x = r.x, y = r.y, z = r.z, wNorm = r.w;
      gid = oids[o0 + ol];
    }
    double acc[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      acc[cc] = 0.0;
    if (STORED) {
      if (have) { // out_i = sum_j A(i, j) lambda_j with the stored column j at ap[j * ld] (coalesced over the lanes)
        const int     ld = (m + 1) & ~1;
        const double *ap = args.eval + args.evalOff[c] + o0 + ol;
        const double *vp = vt + 3;
        int           j  = 0;
        double        acc2[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          acc2[cc] = 0.0;
        for (; j + 8 <= n; j += 8, ap += 8 * static_cast<long long>(ld), vp += 8 * LDV) {
          double a[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            a[u] = __ldg(ap + u * static_cast<long long>(ld));
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
#pragma unroll
            for (int cc = 0; cc < NC; ++cc) {
              acc[cc]  = fma(a[u], vp[u * LDV + cc], acc[cc]);
              acc2[cc] = fma(a[u + 1], vp[(u + 1) * LDV + cc], acc2[cc]);
            }
          }
        }
        for (; j < n; ++j, ap += ld, vp += LDV) {
          const double a = __ldg(ap);
#pragma unroll
          for (int cc = 0; cc < NC; ++cc)
            acc[cc] = fma(a, vp[cc], acc[cc]);
        }
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          acc[cc] += acc2[cc];
      }
    } else if (have) {
      const double *vp = vt + sub * LDV;
      const int     stride = split * LDV;
      int j = sub;
      constexpr int kBatch = 4; // in-vertices evaluated side by side (see pairBasisBatch)
      for (; j + (kBatch - 1) * split < n; j += kBatch * split, vp += kBatch * stride) {
        double bx[kBatch], by[kBatch], bz[kBatch], cf[kBatch][NC], phi[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
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
        pairBasisBatch<KIND, DIM3, kBatch>(x, y, z, bx, by, bz, rbf, invR2, phi);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double part[kBatch / 2];
#pragma unroll
          for (int u = 0; u < kBatch / 2; ++u)
            part[u] = fma(phi[2 * u], cf[2 * u][cc], phi[2 * u + 1] * cf[2 * u + 1][cc]);
#pragma unroll
          for (int u = 1; u < kBatch / 2; ++u)
            part[0] += part[u];
          acc[cc] += part[0];
        }
      }
      for (; j < n; j += split, vp += stride) {
        const double2 xy  = *reinterpret_cast<const double2 *>(vp);
        const double2 zb  = *reinterpret_cast<const double2 *>(vp + 2);
        const double  phi = pairBasis<KIND, DIM3>(x, y, z, xy.x, xy.y, zb.x, rbf, invR2);
        acc[0]            = fma(phi, zb.y, acc[0]);
        if (NC > 1) {
          const double2 b12 = *reinterpret_cast<const double2 *>(vp + 4);
          acc[1]            = fma(phi, b12.x, acc[1]);
          if (NC > 2)
            acc[2] = fma(phi, b12.y, acc[2]);
        }
      }
    }
    if (split > 1) { // warp-uniform
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
      const double w = wNorm;
      if (args.blendStage) { // deterministic blend: summed per vertex in cluster order afterwards (pumBlendGather)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          args.blendStage[(en.outOff + o0 + ol) * dims + c0 + cc] = acc[cc] * w;
      } else {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          atomicAdd(&out[gid * dims + c0 + cc], acc[cc] * w);
      }
    }
  }
  RBF_STAMP(3);
#ifdef RBF_PHASE_TIMING
  if (threadIdx.x == 0)
    atomicAdd(&g_mapPhase[4], 1ull);
#endif
}

template <int KIND, int NC, int NT, bool STAGE_M>
void launchConsistent(const PumSolveArgs &a, const PumEntry *list, int64_t count, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
{
  const size_t tri = STAGE_M ? static_cast<size_t>(pumMatSize(maxN, a.inverse)) : 0;
  const size_t sm  = (tri + static_cast<size_t>(TileLd<NC>::value) * maxN + 16 * (NT / 32)) * sizeof(double);
  if (sm > kPumMaxSmem)
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size.");
  if (a.eval) { // execution-mode minimal-compute
    allowSmem(pumMapConsistentKernel<KIND_RUNTIME, NC, NT, STAGE_M, true, true>, sm);
    pumMapConsistentKernel<KIND_RUNTIME, NC, NT, STAGE_M, true, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else if (a.dim == 3) {
    allowSmem(pumMapConsistentKernel<KIND, NC, NT, STAGE_M, true>, sm);
    pumMapConsistentKernel<KIND, NC, NT, STAGE_M, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else {
    allowSmem(pumMapConsistentKernel<KIND, NC, NT, STAGE_M, false>, sm);
    pumMapConsistentKernel<KIND, NC, NT, STAGE_M, false><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  }
  RBF_LAUNCHED();
}

template <int KIND, int NC>
void bucketConsistent(const PumSolveArgs &a, int bucket, const PumEntry *list, int64_t count, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
{
  if (bucket < 8)
    launchConsistent<KIND, NC, 32, true>(a, list, count, maxN, in, dims, c0, out, s);
  else if (bucket < 12)
    launchConsistent<KIND, NC, 128, true>(a, list, count, maxN, in, dims, c0, out, s);
  else
    launchConsistent<KIND, NC, 256, false>(a, list, count, maxN, in, dims, c0, out, s);
}

// sizes per launch group of the warp-per-cluster buckets (RBFMAP_MAP_SIZE_GROUP overrides, for experiments)
static const int kMapSizeGroup = [] {
  const char *e = std::getenv("RBFMAP_MAP_SIZE_GROUP");
  const int   v = e ? std::atoi(e) : 4;
  return v >= 1 && v <= 8 ? v : 4;
}();

template <int KIND>
void dispatchConsistent(const PumSolveArgs &a, const PumBuckets &bk, const double *in, int dims, double *out, cudaStream_t s)
{
  for (int b = 0; b < kNumBuckets; ++b) {
    const int64_t count = bk.start[b + 1] - bk.start[b];
    if (count <= 0)
      continue;
    for (int c0 = 0; c0 < dims;) { // components are processed in groups of up to 3 per launch
      const int nc = std::min(3, dims - c0);
      // one launch per group of cluster sizes, with the shared memory (and hence the residency) of exactly those sizes
      pumForSizeGroups(bk, b, b < 8 ? kMapSizeGroup : 4, [&](int64_t first, int64_t cnt, int maxN) {
        const PumEntry *list = bk.entries.p + first;
        if (nc == 1)
          bucketConsistent<KIND, 1>(a, b, list, cnt, maxN, in, dims, c0, out, s);
        else if (nc == 2)
          bucketConsistent<KIND, 2>(a, b, list, cnt, maxN, in, dims, c0, out, s);
        else
          bucketConsistent<KIND, 3>(a, b, list, cnt, maxN, in, dims, c0, out, s);
      });
      c0 += nc;
    }
  }
}

} // namespace

void pumMapConsistent(const PumSolveArgs &a, const PumBuckets &bk, const double *in, int dims, double *out, cudaStream_t s)
{
#ifdef RBF_PHASE_TIMING
  unsigned long long zero[8] = {0};
  RBF_CUDA(cudaMemcpyToSymbolAsync(g_mapPhase, zero, sizeof(zero), 0, cudaMemcpyHostToDevice, s));
#endif
  RBF_DISPATCH_HOT_BASIS(a.rbf.kind, dispatchConsistent<KIND>(a, bk, in, dims, out, s));
  RBF_CUDA(cudaGetLastError());
#ifdef RBF_PHASE_TIMING
  unsigned long long h[8];
  RBF_CUDA(cudaMemcpyFromSymbolAsync(h, g_mapPhase, sizeof(h), 0, cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  if (h[4])
    fprintf(stderr, "[map phases] cycles per cluster: gather %.0f poly %.0f solve %.0f eval %.0f (clusters %llu)\n", double(h[0]) / h[4], double(h[1]) / h[4],
            double(h[2]) / h[4], double(h[3]) / h[4], h[4]);
#endif
}

} // namespace rbfmap
