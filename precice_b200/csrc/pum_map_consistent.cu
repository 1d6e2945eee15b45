// pum_map_consistent.cu — per-cluster consistent map: gather, separated polynomial, solve, on-the-fly evaluation and
// partition-of-unity blending in one kernel (sm_100a).
//
// Replaces, per cluster, SphericalVertexCluster::mapConsistent (impl/SphericalVertexCluster.hpp:296-335) with
// RadialBasisFctSolver::solveConsistent (RadialBasisFctSolver.hpp:441-468) and the evaluation matrix of buildMatrixA
// (:209-242), which is never stored: A(i,j) = phi(|out_i - in_j|) is re-evaluated from the vertex tile in shared memory.
// (Kokkos prior art: do_batched_solve, device/KokkosPUMKernels_Impl.hpp:491-814.)
#include "pum_map.cuh"

namespace rbfmap {
namespace {

// Shared-memory layout of the map kernels (doubles):
//   mp   : n(n+1)/2   packed factor (only when it fits, STAGE_M)
//   sx,sy,sz : n each
//   b    : NC * n      right-hand sides / coefficients
//   red  : 16 * nWarps reduction scratch
//   (conservative) ox,oy,oz,ou[NC] : chunk of out-vertices, NT each
template <int KIND, int NC, int NT, bool STAGE_M, bool DIM3>
__global__ void __launch_bounds__(NT) pumMapConsistentKernel(PumSolveArgs args, const int *__restrict__ list, const double *__restrict__ in, int dims, int c0,
                                                             double *__restrict__ out)
{
  extern __shared__ __align__(16) double dyn[];
  const int       c    = list[blockIdx.x];
  const int       n    = args.nIn[c];
  const int       m    = args.nOut[c];
  const int       tid  = threadIdx.x;
  const int      *ids  = args.inIds + args.inOff[c];
  const int      *oids = args.outIds + args.outOff[c];
  const double   *mg   = args.mat + args.matOff[c];
  const RbfParams rbf  = args.rbf;
  const PolyRec   rec  = args.poly[c];
  const double    cx = args.centers[3 * c], cy = args.centers[3 * c + 1], cz = args.centers[3 * c + 2];
  const double    rinv = 1.0 / args.radius;
  const long long tri  = static_cast<long long>(n) * (n + 1) / 2;

  double *ms  = dyn;
  double *sx  = STAGE_M ? dyn + tri : dyn;
  double *sy  = sx + n;
  double *sz  = sy + n;
  double *b   = sz + n;
  double *red = b + NC * n;

  if (m == 0)
    return; // nothing to evaluate (uniform per CTA)
  if (STAGE_M)
    for (long long e = tid; e < tri; e += NT)
      ms[e] = mg[e];
  for (int i = tid; i < n; i += NT) {
    const long long id = ids[i];
    const double   *p  = args.inXyz + 3 * id;
    sx[i]              = p[0];
    sy[i]              = p[1];
    sz[i]              = p[2];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      b[cc * n + i] = in[id * dims + c0 + cc]; // SphericalVertexCluster.hpp:312-320
  }
  __syncthreads();
  const double *mp = STAGE_M ? ms : mg;

  // ---- separated polynomial: beta = lstsq(Q, f); f <- f - Q beta (RadialBasisFctSolver.hpp:446-449) ----
  double beta[4][NC];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      beta[q][cc] = 0.0;
  if (rec.mode == 1) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) { // normal equations + one refinement step on the residual
      double t[4 * NC];
#pragma unroll
      for (int e = 0; e < 4 * NC; ++e)
        t[e] = 0.0;
      for (int i = tid; i < n; i += NT) {
        double q[4];
        polyRow(rec.mode, rec.mask, sx[i], sy[i], sz[i], cx, cy, cz, rinv, q);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double r = b[cc * n + i];
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
            s += rec.G[4 * k + l] * t[l * NC + cc];
          beta[k][cc] += s;
        }
    }
  } else if (rec.mode == 2) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        double s = 0.0;
        for (int i = 0; i < n && i < 3; ++i)
          s += rec.G[4 * k + i] * b[cc * n + i];
        beta[k][cc] = s;
      }
  }
  if (rec.mode != 0) {
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      double q[4];
      polyRow(rec.mode, rec.mask, sx[i], sy[i], sz[i], cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        double r = b[cc * n + i];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          r -= q[k] * beta[k][cc];
        b[cc * n + i] = r;
      }
    }
    __syncthreads();
  }

  // ---- lambda = C^-1 f ----
  blockedSolve<NC>(mp, n, b, n, tid, NT);

  // ---- out = A lambda + V beta, weighted accumulation (SphericalVertexCluster.hpp:322-334) ----
  for (int o = tid; o < m; o += NT) {
    const long long gid = oids[o];
    const double   *p   = args.outXyz + 3 * gid;
    const double    x = p[0], y = p[1], z = p[2];
    double          acc[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      acc[cc] = 0.0;
    for (int j = 0; j < n; ++j) {
      const double d2  = sqDistD<DIM3>(x, y, z, sx[j], sy[j], sz[j]);
      const double phi = evalBasisK<KIND>(distSqrt(d2), rbf);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        acc[cc] = fma(phi, b[cc * n + j], acc[cc]);
    }
    if (rec.mode != 0) {
      double q[4];
      polyRow(rec.mode, rec.mask, x, y, z, cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          acc[cc] += q[k] * beta[k][cc];
    }
    const double w = normalizedWeight(x, y, z, cx, cy, cz, rinv, args.wsum[gid], args.wcount[gid]);
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      atomicAdd(&out[gid * dims + c0 + cc], acc[cc] * w);
  }
}


template <int KIND, int NC, int NT, bool STAGE_M>
void launchConsistent(const PumSolveArgs &a, const int *list, int64_t count, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
{
  const size_t tri = STAGE_M ? static_cast<size_t>(maxN) * (maxN + 1) / 2 : 0;
  const size_t sm  = (tri + (3 + NC) * static_cast<size_t>(maxN) + 16 * (NT / 32) + 0) * sizeof(double);
  if (sm > kPumMaxSmem)
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size.");
  if (a.dim == 3) {
    allowSmem(pumMapConsistentKernel<KIND, NC, NT, STAGE_M, true>, sm);
    pumMapConsistentKernel<KIND, NC, NT, STAGE_M, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else {
    allowSmem(pumMapConsistentKernel<KIND, NC, NT, STAGE_M, false>, sm);
    pumMapConsistentKernel<KIND, NC, NT, STAGE_M, false><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  }
  RBF_LAUNCHED();
}

template <int KIND, int NC>
void bucketConsistent(const PumSolveArgs &a, int bucket, const int *list, int64_t count, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
{
  if (bucket < 8)
    launchConsistent<KIND, NC, 64, true>(a, list, count, maxN, in, dims, c0, out, s);
  else if (bucket < 12)
    launchConsistent<KIND, NC, 128, true>(a, list, count, maxN, in, dims, c0, out, s);
  else
    launchConsistent<KIND, NC, 256, false>(a, list, count, maxN, in, dims, c0, out, s);
}

template <int KIND>
void dispatchConsistent(const PumSolveArgs &a, const PumBuckets &bk, const double *in, int dims, double *out, cudaStream_t s)
{
  for (int b = 0; b < kNumBuckets; ++b) {
    const int64_t count = bk.start[b + 1] - bk.start[b];
    if (count <= 0)
      continue;
    const int *list = bk.order.p + bk.start[b];
    for (int c0 = 0; c0 < dims;) { // components are processed in groups of up to 3 per launch
      const int nc = std::min(3, dims - c0);
      if (nc == 1)
        bucketConsistent<KIND, 1>(a, b, list, count, bk.maxN[b], in, dims, c0, out, s);
      else if (nc == 2)
        bucketConsistent<KIND, 2>(a, b, list, count, bk.maxN[b], in, dims, c0, out, s);
      else
        bucketConsistent<KIND, 3>(a, b, list, count, bk.maxN[b], in, dims, c0, out, s);
      c0 += nc;
    }
  }
}

} // namespace

void pumMapConsistent(const PumSolveArgs &a, const PumBuckets &bk, const double *in, int dims, double *out, cudaStream_t s)
{
  RBF_DISPATCH_HOT_BASIS(a.rbf.kind, dispatchConsistent<KIND>(a, bk, in, dims, out, s));
  RBF_CUDA(cudaGetLastError());
}

} // namespace rbfmap
