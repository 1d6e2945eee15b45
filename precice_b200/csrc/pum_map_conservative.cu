// pum_map_conservative.cu — per-cluster conservative map: weighted gather, A^T u on the fly, solve, polynomial
// correction and accumulation in one kernel (sm_100a).
//
// Replaces, per cluster, SphericalVertexCluster::mapConservative (impl/SphericalVertexCluster.hpp:201-240) with
// RadialBasisFctSolver::solveConservative (RadialBasisFctSolver.hpp:413-438) and the evaluation matrix of buildMatrixA
// (:209-242), which is never stored: A(i,j) = phi(|out_i - in_j|) is re-evaluated from the vertex tiles in shared memory.
// (The reference has no device path for this direction: PartitionOfUnityMapping.hpp:354-355.)
#include "pum_map.cuh"

namespace rbfmap {
namespace {

#ifndef RBF_CONS_BATCH
#define RBF_CONS_BATCH 4
#endif
#ifndef RBF_CONS_MINBLOCKS
#define RBF_CONS_MINBLOCKS 14
#endif
constexpr int kConsMinBlocks = RBF_CONS_MINBLOCKS; // resident clusters per SM the 64-thread kernels are compiled for

// STORED: execution-mode minimal-compute, A is read from args.eval (pum_eval.cu; lanes = in-vertices: entry (i, j) at
// evalOff[c] + i * ld + j) instead of being re-evaluated
template <int KIND, int NC, int NT, bool STAGE_M, bool DIM3, bool STORED = false>
__global__ void __launch_bounds__(NT, NT == 64 ? kConsMinBlocks : 1) pumMapConservativeKernel(PumSolveArgs args, const PumEntry *__restrict__ list, const double *__restrict__ in, int dims, int c0,
                                                               double *__restrict__ out)
{
  constexpr int LDV = TileLd<NC>::value;
  extern __shared__ __align__(16) double dyn[];
  __shared__ __align__(8) unsigned long long mbar;

  const PumEntry en = list[blockIdx.x]; // sizes and offsets in one record: no list -> sizes / offsets -> data chain
  const int c   = en.c;
  const int n   = en.n;
  const int m   = en.m;
  const int tid = threadIdx.x;
  // completeJustInTimeMapping (PartitionOfUnityMapping.hpp:411-422): Au and epsilon come from the cache
  // (SphericalVertexCluster::evaluateConservativeCache :242-253, RadialBasisFctSolver::evaluateConservativeCache :562-573)
  const bool fromCache = args.cacheAu != nullptr;
  if (m == 0 && !fromCache)
    return;
  const long long inOffC = en.inOff;
  const int      *ids  = args.inIds + inOffC;
  const int      *oids = args.outIds + en.outOff;
  const double   *mg   = args.mat + en.matOff;
  const RbfParams rbf  = args.rbf;
  const double    cx = args.centers[3 * c], cy = args.centers[3 * c + 1], cz = args.centers[3 * c + 2];
  const double    rinv = 1.0 / args.radius;
  const double    invR2 = rbf.p1 * rbf.p1;
  const int       triPad = STAGE_M ? static_cast<int>(pumMatSize(n, args.inverse)) : 0;

  double *ms  = dyn;
  double *vt  = dyn + triPad;       // in-vertices: x, y, z, Au / mu
  double *ot  = vt + n * LDV;       // chunk of out-vertices: x, y, z, w*u   (NT rows)
  double *red = ot + NT * LDV;

  if (STAGE_M) {
    if (tid == 0)
      mbarInit(&mbar, 1);
    __syncthreads();
    if (tid == 0)
      bulkLoad(ms, mg, static_cast<unsigned>(triPad) * 8u, &mbar);
  }
  if (STORED && tid == 0) // stored evaluation matrix: on its way into L2 while the tiles are gathered
    bulkPrefetchL2(args.eval + args.evalOff[c], static_cast<unsigned>(((n + 1) & ~1) * m) * 8u);
  for (int i = tid; i < n; i += NT) {
    const double *p = args.inRec + 3 * (inOffC + i); // cluster-ordered coordinates
    vt[i * LDV]     = p[0];
    vt[i * LDV + 1] = p[1];
    vt[i * LDV + 2] = p[2];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      vt[i * LDV + 3 + cc] = fromCache ? args.cacheAu[(inOffC + i) * dims + c0 + cc] : 0.0;
  }
  const int polyMode = args.poly[c].mode, polyMask = args.poly[c].mask, polyRefine = args.poly[c].pad;

  // ---- Au = A^T (w .* u), epsilon = V^T (w .* u) (RadialBasisFctSolver.hpp:420, :428) ----
  double eps[4 * NC];
#pragma unroll
  for (int e = 0; e < 4 * NC; ++e)
    eps[e] = fromCache ? args.cachePoly[(static_cast<long long>(c) * 4 + e / NC) * dims + c0 + e % NC] : 0.0;
  const int rowsPerThread = (n + NT - 1) / NT; // 1 for the register-tiled buckets
  const int mLoop         = fromCache ? 0 : m;
  for (int o0 = 0; o0 < mLoop; o0 += NT) {
    __syncthreads();
    const int o = o0 + tid;
    if (o < m) {
      const long long gid = oids[o];
      const double4   rec = args.outRec[en.outOff + o]; // x, y, z, normalised weight
      const double    x = rec.x, y = rec.y, z = rec.z;
      ot[tid * LDV]     = x;
      ot[tid * LDV + 1] = y;
      ot[tid * LDV + 2] = z;
      // SphericalVertexCluster.hpp:219-227: in(i,c) = data * normalizedWeight
      const double w = rec.w;
      double       q[4];
      if (polyMode != 0)
        polyRow(polyMode, polyMask, x, y, z, cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        const double u          = in[gid * dims + c0 + cc] * w;
        ot[tid * LDV + 3 + cc] = u;
        if (polyMode != 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            eps[k * NC + cc] += q[k] * u;
        }
      }
    }
    __syncthreads();
    const int chunk = min(NT, m - o0);
    for (int r = 0; r < rowsPerThread; ++r) {
      const int j = tid + r * NT;
      if (j < n) {
        double acc[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          acc[cc] = vt[j * LDV + 3 + cc];
        const double  xj = vt[j * LDV], yj = vt[j * LDV + 1], zj = vt[j * LDV + 2];
        const double *op = ot;
        if (STORED) {
          const int     ld = (n + 1) & ~1;
          const double *ap = args.eval + args.evalOff[c] + static_cast<long long>(o0) * ld + j;
          int           oo = 0;
          for (; oo + 4 <= chunk; oo += 4, op += 4 * LDV, ap += 4 * static_cast<long long>(ld)) {
            double a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              a[u] = __ldg(ap + u * static_cast<long long>(ld));
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int cc = 0; cc < NC; ++cc)
                acc[cc] = fma(a[u], op[u * LDV + 3 + cc], acc[cc]);
          }
          for (; oo < chunk; ++oo, op += LDV, ap += ld) {
            const double a = __ldg(ap);
#pragma unroll
            for (int cc = 0; cc < NC; ++cc)
              acc[cc] = fma(a, op[3 + cc], acc[cc]);
          }
        } else {
          // four out-vertices side by side (pairBasisBatch: the dependent FP64 chain of one evaluation uses a fraction of the pipe)
          constexpr int kBatch = RBF_CONS_BATCH;
          int           oo     = 0;
          for (; oo + kBatch <= chunk; oo += kBatch, op += kBatch * LDV) {
            double bx[kBatch], by[kBatch], bz[kBatch], cf[kBatch][NC], phi[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              const double2 xy = *reinterpret_cast<const double2 *>(op + u * LDV);
              const double2 zu = *reinterpret_cast<const double2 *>(op + u * LDV + 2);
              bx[u] = xy.x, by[u] = xy.y, bz[u] = zu.x, cf[u][0] = zu.y;
              if (NC > 1) {
                const double2 u12 = *reinterpret_cast<const double2 *>(op + u * LDV + 4);
                cf[u][1]          = u12.x;
                if (NC > 2)
                  cf[u][2] = u12.y;
              }
            }
            pairBasisBatch<KIND, DIM3, kBatch>(xj, yj, zj, bx, by, bz, rbf, invR2, phi);
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
#pragma unroll
              for (int cc = 0; cc < NC; ++cc)
                acc[cc] = fma(phi[u], cf[u][cc], acc[cc]);
          }
        for (; oo < chunk; ++oo, op += LDV) {
          const double2 xy  = *reinterpret_cast<const double2 *>(op);
          const double2 zu  = *reinterpret_cast<const double2 *>(op + 2);
          const double  phi = pairBasis<KIND, DIM3>(xy.x, xy.y, zu.x, xj, yj, zj, rbf, invR2);
          acc[0]            = fma(phi, zu.y, acc[0]);
          if (NC > 1) {
            const double2 u12 = *reinterpret_cast<const double2 *>(op + 4);
            acc[1]            = fma(phi, u12.x, acc[1]);
            if (NC > 2)
              acc[2] = fma(phi, u12.y, acc[2]);
          }
        }
        }
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          vt[j * LDV + 3 + cc] = acc[cc];
      }
    }
  }
  if (polyMode != 0 && !fromCache)
    blockReduceSum<4 * NC>(eps, red, NT);
  __syncthreads();

  // ---- mu = C^-1 Au ----
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

  // ---- out = mu + Q (Q^T Q)^-1 (epsilon - Q^T mu)  (RadialBasisFctSolver.hpp:427-436) ----
  double gamma[4][NC];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      gamma[k][cc] = 0.0;
  const double *G = args.poly[c].G;
  if (polyMode == 1) {
#pragma unroll 1
    for (int pass = 0; pass <= polyRefine; ++pass) {
      // tau = epsilon - Q^T (mu + Q gamma)
      double t[4 * NC];
#pragma unroll
      for (int e = 0; e < 4 * NC; ++e)
        t[e] = 0.0;
      for (int i = tid; i < n; i += NT) {
        double q[4];
        polyRow(1, polyMask, vt[i * LDV], vt[i * LDV + 1], vt[i * LDV + 2], cx, cy, cz, rinv, q);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double v = vt[i * LDV + 3 + cc];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            v += q[k] * gamma[k][cc];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            t[k * NC + cc] += q[k] * v;
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
            s += G[4 * k + l] * (eps[l * NC + cc] - t[l * NC + cc]);
          gamma[k][cc] += s;
        }
    }
  }
  for (int i = tid; i < n; i += NT) {
    const long long gid = ids[i];
    double          res[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      res[cc] = vt[i * LDV + 3 + cc];
    if (polyMode == 1) {
      double q[4];
      polyRow(1, polyMask, vt[i * LDV], vt[i * LDV + 1], vt[i * LDV + 2], cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          res[cc] += q[k] * gamma[k][cc];
    } else if (polyMode == 2) {
      // y = S^T (epsilon - Q^T mu) with the raw basis; n <= 3 rows
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        double tau[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tau[k] = eps[k * NC + cc];
        for (int r = 0; r < n && r < 3; ++r) {
          double q[4];
          polyRow(2, polyMask, vt[r * LDV], vt[r * LDV + 1], vt[r * LDV + 2], cx, cy, cz, rinv, q);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tau[k] -= q[k] * vt[r * LDV + 3 + cc];
        }
        double yv = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          yv += G[4 * k + i] * tau[k];
        res[cc] += yv;
      }
    }
    // SphericalVertexCluster.hpp:233-239
    if (args.blendStage && !fromCache) { // deterministic blend (pumBlendGather)
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        args.blendStage[(inOffC + i) * dims + c0 + cc] = res[cc];
    } else {
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        atomicAdd(&out[gid * dims + c0 + cc], res[cc]);
    }
  }
}

template <int KIND, int NC, int NT, bool STAGE_M>
void launchConservative(const PumSolveArgs &a, const PumEntry *list, int64_t count, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
{
  const size_t tri = STAGE_M ? static_cast<size_t>(pumMatSize(maxN, a.inverse)) : 0;
  const size_t sm  = (tri + static_cast<size_t>(TileLd<NC>::value) * (maxN + NT) + 16 * (NT / 32)) * sizeof(double);
  if (sm > kPumMaxSmem)
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size.");
  if (a.eval && !a.cacheAu) { // execution-mode minimal-compute
    allowSmem(pumMapConservativeKernel<KIND_RUNTIME, NC, NT, STAGE_M, true, true>, sm);
    pumMapConservativeKernel<KIND_RUNTIME, NC, NT, STAGE_M, true, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else if (a.dim == 3) {
    allowSmem(pumMapConservativeKernel<KIND, NC, NT, STAGE_M, true>, sm);
    pumMapConservativeKernel<KIND, NC, NT, STAGE_M, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else {
    allowSmem(pumMapConservativeKernel<KIND, NC, NT, STAGE_M, false>, sm);
    pumMapConservativeKernel<KIND, NC, NT, STAGE_M, false><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  }
  RBF_LAUNCHED();
}

template <int KIND, int NC>
void bucketConservative(const PumSolveArgs &a, int bucket, const PumEntry *list, int64_t count, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
{
  if (bucket < 8)
    launchConservative<KIND, NC, 64, true>(a, list, count, maxN, in, dims, c0, out, s);
  else if (bucket < 12)
    launchConservative<KIND, NC, 128, true>(a, list, count, maxN, in, dims, c0, out, s);
  else
    launchConservative<KIND, NC, 256, false>(a, list, count, maxN, in, dims, c0, out, s);
}

template <int KIND>
void dispatchConservative(const PumSolveArgs &a, const PumBuckets &bk, const double *in, int dims, double *out, cudaStream_t s)
{
  for (int b = 0; b < kNumBuckets; ++b) {
    const int64_t count = bk.start[b + 1] - bk.start[b];
    if (count <= 0)
      continue;
    for (int c0 = 0; c0 < dims;) {
      const int nc = std::min(3, dims - c0);
      // one launch per group of cluster sizes (see pumForSizeGroups)
      pumForSizeGroups(bk, b, b < 8 ? 2 : 4, [&](int64_t first, int64_t cnt, int maxN) {
        const PumEntry *list = bk.entries.p + first;
        if (nc == 1)
          bucketConservative<KIND, 1>(a, b, list, cnt, maxN, in, dims, c0, out, s);
        else if (nc == 2)
          bucketConservative<KIND, 2>(a, b, list, cnt, maxN, in, dims, c0, out, s);
        else
          bucketConservative<KIND, 3>(a, b, list, cnt, maxN, in, dims, c0, out, s);
      });
      c0 += nc;
    }
  }
}

} // namespace

void pumMapConservative(const PumSolveArgs &a, const PumBuckets &bk, const double *in, int dims, double *out, cudaStream_t s)
{
  RBF_DISPATCH_HOT_BASIS(a.rbf.kind, dispatchConservative<KIND>(a, bk, in, dims, out, s));
  RBF_CUDA(cudaGetLastError());
}

} // namespace rbfmap
