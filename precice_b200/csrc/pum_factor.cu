// pum_factor.cu — per-cluster assembly of the interpolation matrix C and its factorisation (sm_100a).
//
// Replaces, per cluster, RadialBasisFctSolver.hpp:164-207 (buildMatrixCLU) and :352-353 (Eigen::LLT), i.e. the
// "map.pou.computeMapping.rbfSolver" part of SphericalVertexCluster's constructor (SphericalVertexCluster.hpp:178-181)
// and the Kokkos kernels do_input_assembly / do_batched_lu (device/KokkosPUMKernels_Impl.hpp:305-380, 458-485).
//
// One CTA per cluster. The matrix never touches memory before it is factorised: element (i, j) =
// (a*TG + tr, b*TG + tc) lives in register A[a][b] of thread (tr, tc) (block-cyclic distribution, NB x NB blocks
// per thread), so the active trailing part stays evenly spread over the threads while the factorisation proceeds.
// Factorisation: C = M D^-1 M^T, the square-root-free form of the Cholesky factorisation (M = L diag(L_kk),
// D = diag(L_kk^2)); the pivots d_k are exactly the squares of Eigen's LLT pivots, hence "d_k <= 0" is the same
// failure criterion as LLT's info() != Success. Packed result per cluster (column-major lower triangle,
// n(n+1)/2 doubles): column k = [1/d_k, M(k+1,k), ..., M(n-1,k)].
#include "pum_device.cuh"

#include <climits>
#include <cstdlib>

namespace rbfmap {
namespace {

// row length (doubles) of the column exchange buffer: 8 rows 80 B apart hit 8 different 16-byte bank groups,
// so the LDS.128 operand loads of a warp are conflict free
constexpr int kColLd = 10;

template <int KIND, int TG, int NB, bool DIM3>
__global__ void __launch_bounds__(TG *TG, TG == 8 ? (NB <= 7 ? 10 : 8) : 2) pumFactorKernel(PumSolveArgs args, const int *__restrict__ list, int *__restrict__ fail)
{
  constexpr int NT   = TG * TG;
  constexpr int NMAX = NB * TG;
  __shared__ double sx[NMAX], sy[NMAX], sz[NMAX];
  __shared__ __align__(16) double colBuf[2][2][TG * kColLd]; // [parity][unscaled | scaled by 1/d_k]

  const int c   = list[blockIdx.x];
  const int n   = args.nIn[c];
  const int tid = threadIdx.x;
  const int tr = tid % TG, tc = tid / TG;
  const int      *ids = args.inIds + args.inOff[c];
  const RbfParams rbf   = args.rbf;
  const double    invR2 = rbf.p1 * rbf.p1;

  for (int i = tid; i < NMAX; i += NT) {
    double x = paddingCoordinate(i), y = 0.0, z = 0.0;
    if (i < n) {
      const double *p = args.inXyz + 3 * static_cast<long long>(ids[i]);
      x = p[0];
      y = p[1];
      z = p[2];
    }
    sx[i] = x;
    sy[i] = y;
    sz[i] = z;
  }
  __syncthreads();

  // ---- assembly of the lower block triangle (buildMatrixCLU): phi(|x_i - x_j|) ----
  double A[NB][NB];
  {
    double xj[NB], yj[NB], zj[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      xj[b] = sx[b * TG + tc];
      yj[b] = sy[b * TG + tc];
      zj[b] = sz[b * TG + tc];
    }
#pragma unroll
    for (int a = 0; a < NB; ++a) {
      const double xi = sx[a * TG + tr], yi = sy[a * TG + tr], zi = sz[a * TG + tr];
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        // RadialBasisFctSolver.hpp:191-192
        A[a][b] = pairBasis<KIND, DIM3>(xj[b], yj[b], zj[b], xi, yi, zi, rbf, invR2);
      }
    }
  }

  // ---- right-looking factorisation, one column per step, one barrier per step ----
  // Column k is published in shared memory twice: unscaled (row operand M(i,k)) and scaled by 1/d_k (column operand).
  // Look-ahead: in step k every thread first updates its entries of the block column that holds column k+1; the warp that
  // owns column k+1 (its TG owners are consecutive lanes of one warp) then fetches the new pivot from the diagonal thread
  // with a shuffle, inverts it and publishes column k+1 into the other buffer while all warps finish the trailing update.
  // All NB*TG columns are processed: the padding vertices form an identity block that costs a few cheap steps and
  // removes every size-dependent branch from the loop. Pivot test: the smallest high word of the pivots is tracked and
  // checked once at the end (d <= 0 <=> high word <= 0, the criterion of Eigen::LLT; NaNs pass there as well).
  constexpr int kParity = 2 * TG * kColLd; // doubles between the two buffer pairs
  constexpr int kScaled = TG * kColLd;
  double       *rowP    = &colBuf[0][0][0] + tr * kColLd; // row operand of this thread / publishing slot of an owner
  const double *colP    = &colBuf[0][1][0] + tc * kColLd; // column operand
  const int     warp    = tid >> 5;
  int           minHi   = 0x7ff00000;

  auto publish = [&](const int KB, const int kk2) { // warp-uniform: executed by all lanes of the warp that owns column kk2
    const int    lane0 = (kk2 * TG) & 31;
    const double d     = __shfl_sync(0xffffffffu, A[KB][KB], lane0 + kk2);
    minHi              = min(minHi, __double2hiint(d));
    const double rd    = pivotReciprocal(d);
    if (tc == kk2) {
      double *dst = rowP + (kk2 & 1) * kParity;
#pragma unroll
      for (int a2 = (KB & ~1); a2 < NB; a2 += 2) {
        const double v0 = a2 >= KB ? A[a2 >= KB ? a2 : KB][KB] : 0.0;
        const double v1 = a2 + 1 < NB ? A[a2 + 1 < NB ? a2 + 1 : KB][KB] : 0.0;
        *reinterpret_cast<double2 *>(dst + a2)           = make_double2(v0, v1);
        *reinterpret_cast<double2 *>(dst + kScaled + a2) = make_double2(v0 * rd, v1 * rd);
      }
      if (tr == kk2)
        A[KB][KB] = rd; // the packed factor stores 1/d_k; this entry is final
    }
  };

  if (warp == 0)
    publish(0, 0);
#pragma unroll
  for (int kb = 0; kb < NB; ++kb) {
    // steps 0 .. TG-2 of the block: the next column lies in the same block column
#pragma unroll 1
    for (int kk = 0; kk < TG - 1; ++kk) {
      const int par = (kk & 1) * kParity;
      __syncthreads();
      double li[NB + 1], lj[NB + 1];
#pragma unroll
      for (int a2 = (kb & ~1); a2 < NB; a2 += 2) {
        const double2 u = *reinterpret_cast<const double2 *>(rowP + par + a2);
        const double2 v = *reinterpret_cast<const double2 *>(colP + par + a2);
        li[a2]     = u.x;
        li[a2 + 1] = u.y;
        lj[a2]     = v.x;
        lj[a2 + 1] = v.y;
      }
      const double ljk = tc > kk ? lj[kb] : 0.0; // columns j <= k of block column kb are final
#pragma unroll
      for (int a = kb; a < NB; ++a)
        A[a][kb] = fma(-li[a], ljk, A[a][kb]);
      if (warp == ((kk + 1) * TG) >> 5)
        publish(kb, kk + 1);
#pragma unroll
      for (int a = kb + 1; a < NB; ++a) {
#pragma unroll
        for (int b = kb + 1; b <= a; ++b)
          A[a][b] = fma(-li[a], lj[b], A[a][b]);
      }
    }
    // last step of the block: block column kb is final, the next column is the first one of block column kb + 1
    {
      constexpr int par = ((TG - 1) & 1) * kParity;
      __syncthreads();
      double li[NB + 1], lj[NB + 1];
#pragma unroll
      for (int a2 = (kb & ~1); a2 < NB; a2 += 2) {
        const double2 u = *reinterpret_cast<const double2 *>(rowP + par + a2);
        const double2 v = *reinterpret_cast<const double2 *>(colP + par + a2);
        li[a2]     = u.x;
        li[a2 + 1] = u.y;
        lj[a2]     = v.x;
        lj[a2 + 1] = v.y;
      }
      if (kb + 1 < NB) {
#pragma unroll
        for (int a = kb + 1; a < NB; ++a)
          A[a][kb + 1] = fma(-li[a], lj[kb + 1], A[a][kb + 1]);
        if (warp == 0)
          publish(kb + 1, 0);
#pragma unroll
        for (int a = kb + 2; a < NB; ++a) {
#pragma unroll
          for (int b = kb + 2; b <= a; ++b)
            A[a][b] = fma(-li[a], lj[b], A[a][b]);
        }
      }
    }
  }
  if (__syncthreads_or(minHi <= 0)) {
    if (tid == 0)
      atomicMin(fail, c);
    return;
  }

  // ---- packed store: the owners still hold the unscaled columns M(:,k) and the pivots d_k ----
  double *mp = args.mat + args.matOff[c];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int j = b * TG + tc;
    if (j < n) {
      double *colp = mp + colStart32(j, n) - j;
#pragma unroll
      for (int a = b; a < NB; ++a) {
        const int i = a * TG + tr;
        if (i < n && i >= j)
          colp[i] = A[a][b]; // the diagonal already holds 1/d_j
      }
    }
  }
}

// ---- generic variant for clusters beyond the register-tiled classes: factor in place in the packed global storage ----
template <int KIND>
__global__ void __launch_bounds__(256) pumFactorGenericKernel(PumSolveArgs args, const int *__restrict__ list, int *__restrict__ fail)
{
  extern __shared__ double dyn[]; // sx, sy, sz, col : 4 n doubles
  const int       c   = list[blockIdx.x];
  const int       n   = args.nIn[c];
  const int      *ids = args.inIds + args.inOff[c];
  double         *mp  = args.mat + args.matOff[c];
  const RbfParams rbf = args.rbf;
  const double    invR2 = rbf.p1 * rbf.p1;
  double         *sx = dyn, *sy = dyn + n, *sz = dyn + 2 * n, *col = dyn + 3 * n;
  __shared__ int  bad;
  if (threadIdx.x == 0)
    bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double *p = args.inXyz + 3 * static_cast<long long>(ids[i]);
    sx[i]           = p[0];
    sy[i]           = p[1];
    sz[i]           = p[2];
  }
  __syncthreads();
  const bool dim3 = args.dim == 3;
  for (int j = 0; j < n; ++j) {
    const long long cs = colStart(j, n);
    for (int i = j + threadIdx.x; i < n; i += blockDim.x) {
      mp[cs + i - j] = dim3 ? pairBasis<KIND, true>(sx[j], sy[j], sz[j], sx[i], sy[i], sz[i], rbf, invR2)
                            : pairBasis<KIND, false>(sx[j], sy[j], sz[j], sx[i], sy[i], sz[i], rbf, invR2);
    }
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const long long ck = colStart(k, n);
    for (int i = k + threadIdx.x; i < n; i += blockDim.x)
      col[i] = mp[ck + i - k];
    __syncthreads();
    const double d = col[k];
    if (!(d > 0.0)) {
      if (threadIdx.x == 0)
        bad = 1;
      break;
    }
    const double rd = 1.0 / d;
    if (threadIdx.x == 0)
      mp[ck] = rd;
    // trailing update A(i,j) -= M(i,k) M(j,k) / d for j > k, i >= j: one column per warp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = k + 1 + w; j < n; j += nw) {
      const double    ljk = col[j] * rd;
      double         *t   = mp + colStart(j, n) - j;
      for (int i = j + lane; i < n; i += 32)
        t[i] = fma(-col[i], ljk, t[i]);
    }
    __syncthreads();
  }
  __syncthreads();
  if (bad && threadIdx.x == 0)
    atomicMin(fail, c);
}

template <int KIND, int TG, int NB>
void launchTiled(const PumSolveArgs &a, const int *list, int64_t count, int *fail, cudaStream_t s)
{
  if (a.dim == 3)
    pumFactorKernel<KIND, TG, NB, true><<<static_cast<unsigned>(count), TG * TG, 0, s>>>(a, list, fail);
  else
    pumFactorKernel<KIND, TG, NB, false><<<static_cast<unsigned>(count), TG * TG, 0, s>>>(a, list, fail);
}

template <int KIND>
void factorBucket(const PumSolveArgs &a, int bucket, const int *list, int64_t count, int maxN, int *fail, cudaStream_t s)
{
  switch (bucket) {
  case 0: launchTiled<KIND, 8, 1>(a, list, count, fail, s); break;
  case 1: launchTiled<KIND, 8, 2>(a, list, count, fail, s); break;
  case 2: launchTiled<KIND, 8, 3>(a, list, count, fail, s); break;
  case 3: launchTiled<KIND, 8, 4>(a, list, count, fail, s); break;
  case 4: launchTiled<KIND, 8, 5>(a, list, count, fail, s); break;
  case 5: launchTiled<KIND, 8, 6>(a, list, count, fail, s); break;
  case 6: launchTiled<KIND, 8, 7>(a, list, count, fail, s); break;
  case 7: launchTiled<KIND, 8, 8>(a, list, count, fail, s); break;
  case 8: launchTiled<KIND, 16, 5>(a, list, count, fail, s); break;
  case 9: launchTiled<KIND, 16, 6>(a, list, count, fail, s); break;
  case 10: launchTiled<KIND, 16, 7>(a, list, count, fail, s); break;
  case 11: launchTiled<KIND, 16, 8>(a, list, count, fail, s); break;
  default: {
    const size_t sm = 4 * static_cast<size_t>(maxN) * sizeof(double);
    if (sm > kPumMaxSmem)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size.");
    if (sm > 48 * 1024)
      RBF_CUDA(cudaFuncSetAttribute(pumFactorGenericKernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm)));
    pumFactorGenericKernel<KIND><<<static_cast<unsigned>(count), 256, sm, s>>>(a, list, fail);
  }
  }
  RBF_LAUNCHED();
}

} // namespace

// RBFMAP_TILED_FACTOR=1 keeps the column-by-column kernel for all sizes (A/B comparison and tests)
static const bool g_pumForceTiledFactor = [] { const char *e = std::getenv("RBFMAP_TILED_FACTOR"); return e && e[0] == '1'; }();

void pumFactor(const PumSolveArgs &a, const PumBuckets &bk, int *fail, cudaStream_t s)
{
  for (int b = 0; b < kNumBuckets; ++b) {
    const int64_t count = bk.start[b + 1] - bk.start[b];
    if (count <= 0)
      continue;
    const int *list = bk.order.p + bk.start[b];
    if (b < 8 && !g_pumForceTiledFactor) { // clusters of up to 64 vertices: one warp per cluster on DMMA (pum_factor_warp.cu)
      pumFactorWarp(a, b, bk.entries.p + bk.start[b], count, fail, s);
      continue;
    }
    RBF_DISPATCH_HOT_BASIS(a.rbf.kind, factorBucket<KIND>(a, b, list, count, bk.maxN[b], fail, s));
  }
  RBF_CUDA(cudaGetLastError());
}

} // namespace rbfmap
