// pum_map_conservative.cu — per-cluster conservative map: weighted gather, A^T u on the fly, solve, polynomial
// correction and accumulation in one kernel (sm_100a).
//
// Replaces, per cluster, SphericalVertexCluster::mapConservative (impl/SphericalVertexCluster.hpp:201-240) with
// RadialBasisFctSolver::solveConservative (RadialBasisFctSolver.hpp:413-438) and the evaluation matrix of buildMatrixA
// (:209-242), which is never stored: A(i,j) = phi(|out_i - in_j|) is re-evaluated from the vertex tile in shared memory.
// (The reference has no device path for this direction: PartitionOfUnityMapping.hpp:354-355.)
#include "pum_map.cuh"

namespace rbfmap {
namespace {

template <int KIND, int NC, int NT, bool STAGE_M, bool DIM3>
__global__ void __launch_bounds__(NT) pumMapConservativeKernel(PumSolveArgs args, const int *__restrict__ list, const double *__restrict__ in, int dims, int c0,
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
  double *ox  = red + 16 * (NT / 32);
  double *oy  = ox + NT;
  double *oz  = oy + NT;
  double *ou  = oz + NT; // NC * NT

  if (m == 0)
    return;
  if (STAGE_M)
    for (long long e = tid; e < tri; e += NT)
      ms[e] = mg[e];
  for (int i = tid; i < n; i += NT) {
    const double *p = args.inXyz + 3 * static_cast<long long>(ids[i]);
    sx[i]           = p[0];
    sy[i]           = p[1];
    sz[i]           = p[2];
  }
  const double *mp = STAGE_M ? ms : mg;

  // ---- Au = A^T (w .* u), epsilon = V^T (w .* u) (RadialBasisFctSolver.hpp:420, :428) ----
  double eps[4 * NC];
#pragma unroll
  for (int e = 0; e < 4 * NC; ++e)
    eps[e] = 0.0;
  const int rowsPerThread = (n + NT - 1) / NT; // 1 for the register-tiled classes
  for (int i = tid; i < n; i += NT)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      b[cc * n + i] = 0.0;
  for (int o0 = 0; o0 < m; o0 += NT) {
    __syncthreads();
    const int o = o0 + tid;
    if (o < m) {
      const long long gid = oids[o];
      const double   *p   = args.outXyz + 3 * gid;
      const double    x = p[0], y = p[1], z = p[2];
      ox[tid] = x;
      oy[tid] = y;
      oz[tid] = z;
      // SphericalVertexCluster.hpp:219-227: in(i,c) = data * normalizedWeight
      const double w = normalizedWeight(x, y, z, cx, cy, cz, rinv, args.wsum[gid], args.wcount[gid]);
      double       q[4];
      if (rec.mode != 0)
        polyRow(rec.mode, rec.mask, x, y, z, cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        const double u   = in[gid * dims + c0 + cc] * w;
        ou[cc * NT + tid] = u;
        if (rec.mode != 0) {
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
          acc[cc] = b[cc * n + j];
        const double xj = sx[j], yj = sy[j], zj = sz[j];
        for (int oo = 0; oo < chunk; ++oo) {
          const double d2  = sqDistD<DIM3>(ox[oo], oy[oo], oz[oo], xj, yj, zj);
          const double phi = evalBasisK<KIND>(distSqrt(d2), rbf);
#pragma unroll
          for (int cc = 0; cc < NC; ++cc)
            acc[cc] = fma(phi, ou[cc * NT + oo], acc[cc]);
        }
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          b[cc * n + j] = acc[cc];
      }
    }
  }
  if (rec.mode != 0)
    blockReduceSum<4 * NC>(eps, red, NT);
  __syncthreads();

  // ---- mu = C^-1 Au ----
  blockedSolve<NC>(mp, n, b, n, tid, NT);

  // ---- out = mu + Q (Q^T Q)^-1 (epsilon - Q^T mu)  (RadialBasisFctSolver.hpp:427-436) ----
  double gamma[4][NC];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      gamma[k][cc] = 0.0;
  if (rec.mode == 1) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      // tau = epsilon - Q^T (mu + Q gamma)
      double t[4 * NC];
#pragma unroll
      for (int e = 0; e < 4 * NC; ++e)
        t[e] = 0.0;
      for (int i = tid; i < n; i += NT) {
        double q[4];
        polyRow(rec.mode, rec.mask, sx[i], sy[i], sz[i], cx, cy, cz, rinv, q);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double v = b[cc * n + i];
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
            s += rec.G[4 * k + l] * (eps[l * NC + cc] - t[l * NC + cc]);
          gamma[k][cc] += s;
        }
    }
  }
  for (int i = tid; i < n; i += NT) {
    const long long gid = ids[i];
    double          res[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      res[cc] = b[cc * n + i];
    if (rec.mode == 1) {
      double q[4];
      polyRow(rec.mode, rec.mask, sx[i], sy[i], sz[i], cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          res[cc] += q[k] * gamma[k][cc];
    } else if (rec.mode == 2) {
      // y = S^T (epsilon - Q^T mu) with the raw basis; n <= 3 rows
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        double tau[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tau[k] = eps[k * NC + cc];
        for (int r = 0; r < n && r < 3; ++r) {
          double q[4];
          polyRow(rec.mode, rec.mask, sx[r], sy[r], sz[r], cx, cy, cz, rinv, q);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tau[k] -= q[k] * b[cc * n + r];
        }
        double yv = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          yv += rec.G[4 * k + i] * tau[k];
        res[cc] += yv;
      }
    }
    // SphericalVertexCluster.hpp:233-239
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      atomicAdd(&out[gid * dims + c0 + cc], res[cc]);
  }
}


template <int KIND, int NC, int NT, bool STAGE_M>
void launchConservative(const PumSolveArgs &a, const int *list, int64_t count, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
{
  const size_t tri = STAGE_M ? static_cast<size_t>(maxN) * (maxN + 1) / 2 : 0;
  const size_t sm  = (tri + (3 + NC) * static_cast<size_t>(maxN) + 16 * (NT / 32) + (3 + NC) * NT) * sizeof(double);
  if (sm > kPumMaxSmem)
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size.");
  if (a.dim == 3) {
    allowSmem(pumMapConservativeKernel<KIND, NC, NT, STAGE_M, true>, sm);
    pumMapConservativeKernel<KIND, NC, NT, STAGE_M, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else {
    allowSmem(pumMapConservativeKernel<KIND, NC, NT, STAGE_M, false>, sm);
    pumMapConservativeKernel<KIND, NC, NT, STAGE_M, false><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  }
  RBF_LAUNCHED();
}

template <int KIND, int NC>
void bucketConservative(const PumSolveArgs &a, int bucket, const int *list, int64_t count, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
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
    const int *list = bk.order.p + bk.start[b];
    for (int c0 = 0; c0 < dims;) { // components are processed in groups of up to 3 per launch
      const int nc = std::min(3, dims - c0);
      if (nc == 1)
        bucketConservative<KIND, 1>(a, b, list, count, bk.maxN[b], in, dims, c0, out, s);
      else if (nc == 2)
        bucketConservative<KIND, 2>(a, b, list, count, bk.maxN[b], in, dims, c0, out, s);
      else
        bucketConservative<KIND, 3>(a, b, list, count, bk.maxN[b], in, dims, c0, out, s);
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
