// pum_factor_warp.cu — per-cluster assembly + factorisation for clusters of up to 64 vertices: ONE WARP per cluster,
// the matrix held in FP64 tensor-core accumulator fragments, blocked factorisation on DMMA (sm_100a).
//
// Replaces, per cluster, RadialBasisFctSolver.hpp:164-207 (buildMatrixCLU) and :352-353 (Eigen::LLT), like
// pum_factor.cu, which keeps the clusters with more than 64 vertices.
//
// Why one warp: the column-by-column kernel needs a block barrier and a pivot -> reciprocal -> scale -> exchange chain
// per column with only a dozen FMAs per thread in between; it is bound by that latency (ncu: 20 % barrier stalls, FP64
// pipe 35 % busy). Here a warp owns the whole cluster, so there is no block barrier at all, and the O(n^3) part runs
// as 8x8x4 DMMAs (mma.sync.m8n8k4.f64: 256 FMAs per instruction instead of 32).
//
// Layout: the lower block triangle of C in 8x8 blocks; block (a, b) lives in the DMMA accumulator layout: lane
// (g = lane / 4, q = lane % 4) holds elements (8a + g, 8b + 2q) and (8a + g, 8b + 2q + 1).
// Blocked right-looking square-root-free factorisation C = M D^-1 M^T, for each block column kb:
//   1. diagonal block: 8 elimination steps inside the warp (shuffles only). The same row operations applied to the
//      identity give X = Lu^-1, Lu = M_kk D^-1 the unit lower factor of the block.
//   2. panel: M_a = C_a X^T for every block below the diagonal (two DMMAs per block).
//   3. trailing update: C_ab -= M_a D^-1 M_b^T (two DMMAs per block), operands re-read from a per-warp panel buffer
//      in shared memory (the accumulator layout is not the operand layout).
// Packed result exactly as pum_factor.cu writes it: column k = [1/d_k, M(k+1,k), ..., M(n-1,k)].
#include "pum_device.cuh"

#include <climits>

namespace rbfmap {
namespace {

#ifndef RBF_W7
#define RBF_W7 12
#endif
constexpr int kW7 = RBF_W7;
constexpr int kSlotLd = 12; // doubles per row of a panel slot: rows 96 B apart -> conflict-free LDS.64 fragment loads

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// reciprocal of a pivot: MUFU.RCP64H seed y (|e| = |1 - d y| < 2^-20), y (1 + e + e^2): three dependent FMAs
__device__ __forceinline__ double pivotReciprocal3(double d)
{
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  return fma(y, fma(e, e, e), y);
}

template <int KIND, int NB, bool DIM3>
__global__ void __launch_bounds__(32, NB <= 4 ? 20 : NB == 5 ? 16 : NB == 6 ? 14 : NB == 7 ? kW7 : 10) pumFactorWarpKernel(PumSolveArgs args, const PumEntry *__restrict__ list, int *__restrict__ fail)
{
  constexpr int NMAX = NB * 8;
  __shared__ __align__(16) double sx[NMAX], sy[NMAX], sz[NMAX];
  __shared__ __align__(16) double slot[NB][8 * kSlotLd];

  const PumEntry  en   = list[blockIdx.x];
  const int       c    = en.c;
  const int       n    = en.n;
  const int       lane = threadIdx.x;
  const int       g = lane >> 2, q = lane & 3;
  const RbfParams rbf   = args.rbf;
  const double    invR2 = rbf.p1 * rbf.p1;

  for (int i = lane; i < NMAX; i += 32) {
    double x = paddingCoordinate(i), y = 0.0, z = 0.0;
    if (i < n) {
      const double *p = args.inRec + 3 * (en.inOff + i);
      x = p[0];
      y = p[1];
      z = p[2];
    }
    sx[i] = x;
    sy[i] = y;
    sz[i] = z;
  }
  __syncwarp();

  // ---- assembly of the lower block triangle (buildMatrixCLU): phi(|x_i - x_j|), RadialBasisFctSolver.hpp:191-192 ----
  double C[NB][NB][2];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const double2 xj = *reinterpret_cast<const double2 *>(&sx[8 * b + 2 * q]);
    const double2 yj = *reinterpret_cast<const double2 *>(&sy[8 * b + 2 * q]);
    const double2 zj = *reinterpret_cast<const double2 *>(&sz[8 * b + 2 * q]);
#pragma unroll
    for (int a = b; a < NB; ++a) {
      const double xi = sx[8 * a + g], yi = sy[8 * a + g], zi = sz[8 * a + g];
      C[a][b][0] = pairBasis<KIND, DIM3>(xj.x, yj.x, zj.x, xi, yi, zi, rbf, invR2);
      C[a][b][1] = pairBasis<KIND, DIM3>(xj.y, yj.y, zj.y, xi, yi, zi, rbf, invR2);
    }
  }

  int minHi = 0x7ff00000; // smallest high word of the pivots: d <= 0 <=> high word <= 0 (the test of Eigen::LLT)
  double *const myRow  = &slot[0][0] + g * kSlotLd + 2 * q; // accumulator-layout position of this lane inside a slot
  const double *myFrag = &slot[0][0] + g * kSlotLd + q;     // operand-layout position (k = q, + 4 for the second half)
  constexpr int kSlot  = 8 * kSlotLd;

#pragma unroll
  for (int kb = 0; kb < NB; ++kb) {
    // ---- 1. diagonal block: elimination with shuffles; E accumulates the inverse of the unit lower factor ----
    double s0 = C[kb][kb][0], s1 = C[kb][kb][1];
    double e0 = g == 2 * q ? 1.0 : 0.0, e1 = g == 2 * q + 1 ? 1.0 : 0.0;
    double nr0 = 0.0, nr1 = 0.0; // -1/d_k for the two operand columns of this lane: k = q and k = 4 + q
#pragma unroll 1
    for (int kp = 0; kp < 4; ++kp) { // column pair kp: columns 2kp (held in s0 by the lanes with q == kp) and 2kp + 1 (s1)
      // padding columns (index >= n, last block column only) are never stored and only touch padding rows: skip them
      if (kb == NB - 1 && 8 * kb + 2 * kp >= n)
        break;
#pragma unroll
      for (int odd = 0; odd < 2; ++odd) {
        const int    k    = 2 * kp + odd;
        if (kb == NB - 1 && odd == 1 && 8 * kb + k >= n)
          break;
        const double colk = odd ? s1 : s0;
        const int    srcD = 4 * k + kp; // lane that holds S(k, k)
        const double d    = shfl(colk, srcD);
        const double ci   = shfl(colk, (lane & ~3) | kp); // S(g, k)
        const double cj0  = shfl(colk, 8 * q + kp);       // S(2q, k)
        const double cj1  = shfl(colk, 8 * q + 4 + kp);   // S(2q + 1, k)
        const double ek0  = shfl(e0, 4 * k + q);          // E(k, 2q)
        const double ek1  = shfl(e1, 4 * k + q);          // E(k, 2q + 1)
        minHi             = min(minHi, __double2hiint(d));
        const double r    = pivotReciprocal3(d);
        if ((k & 3) == q) {
          if (k < 4)
            nr0 = -r;
          else
            nr1 = -r;
        }
        if (lane == srcD) { // the packed factor stores 1/d_k on the diagonal; this entry is final
          if (odd)
            s1 = r;
          else
            s0 = r;
        }
        if (g > k) { // rows <= k are final
          const double t = ci * r;
          if (2 * q > k) // columns <= k are final
            s0 = fma(-t, cj0, s0);
          if (2 * q + 1 > k)
            s1 = fma(-t, cj1, s1);
          e0 = fma(-t, ek0, e0);
          e1 = fma(-t, ek1, e1);
        }
      }
    }
    C[kb][kb][0] = s0;
    C[kb][kb][1] = s1;

    if (kb + 1 < NB) {
      // ---- 2. panel: M_a = C_a X^T; B operand (k, n) = X(n, k): lane (g, q) needs X(g, 4h + q) ----
      double bx[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int    src = (lane & ~3) | (2 * h + (q >> 1));
        const double v0 = shfl(e0, src), v1 = shfl(e1, src);
        bx[h]           = (q & 1) ? v1 : v0;
      }
#pragma unroll
      for (int a = kb + 1; a < NB; ++a)
        *reinterpret_cast<double2 *>(myRow + a * kSlot) = make_double2(C[a][kb][0], C[a][kb][1]);
      __syncwarp();
#pragma unroll
      for (int a = kb + 1; a < NB; ++a) {
        const double a0 = myFrag[a * kSlot], a1 = myFrag[a * kSlot + 4];
        double       d0 = 0.0, d1 = 0.0;
        dmma(d0, d1, a0, bx[0]);
        dmma(d0, d1, a1, bx[1]);
        C[a][kb][0] = d0;
        C[a][kb][1] = d1;
      }
      __syncwarp();
#pragma unroll
      for (int a = kb + 1; a < NB; ++a)
        *reinterpret_cast<double2 *>(myRow + a * kSlot) = make_double2(C[a][kb][0], C[a][kb][1]);
      __syncwarp();
      // ---- 3. trailing update: C_ab -= M_a D^-1 M_b^T; B operand (k, n) = -M_b(n, k) / d_k ----
#pragma unroll
      for (int b = kb + 1; b < NB; ++b) {
        const double f0 = myFrag[b * kSlot], f1 = myFrag[b * kSlot + 4];
        const double b0 = f0 * nr0, b1 = f1 * nr1;
        dmma(C[b][b][0], C[b][b][1], f0, b0);
        dmma(C[b][b][0], C[b][b][1], f1, b1);
#pragma unroll
        for (int a = b + 1; a < NB; ++a) {
          const double a0 = myFrag[a * kSlot], a1 = myFrag[a * kSlot + 4];
          dmma(C[a][b][0], C[a][b][1], a0, b0);
          dmma(C[a][b][0], C[a][b][1], a1, b1);
        }
      }
      __syncwarp(); // the slots are rewritten by the next panel
    }
  }
  if (minHi <= 0) { // uniform over the warp: every lane saw every pivot
    if (lane == 0)
      atomicMin(fail, c);
    return;
  }

  // ---- packed store ----
  double *mp = args.mat + en.matOff;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int j0 = 8 * b + 2 * q;
    double   *col0 = mp + colStart32(j0, n) - j0;         // column j0: element (i, j0) at col0[i]
    double   *col1 = col0 + (n - j0 - 1);                 // column j0 + 1
#pragma unroll
    for (int a = b; a < NB; ++a) {
      const int i = 8 * a + g;
      if (i < n) {
        if (i >= j0)
          col0[i] = C[a][b][0];
        if (i >= j0 + 1)
          col1[i] = C[a][b][1];
      }
    }
  }
}

template <int KIND, int NB>
void launchWarp(const PumSolveArgs &a, const PumEntry *list, int64_t count, int *fail, cudaStream_t s)
{
  if (a.dim == 3)
    pumFactorWarpKernel<KIND, NB, true><<<static_cast<unsigned>(count), 32, 0, s>>>(a, list, fail);
  else
    pumFactorWarpKernel<KIND, NB, false><<<static_cast<unsigned>(count), 32, 0, s>>>(a, list, fail);
}

template <int KIND>
void warpBucket(const PumSolveArgs &a, int bucket, const PumEntry *list, int64_t count, int *fail, cudaStream_t s)
{
  switch (bucket) {
  case 0: launchWarp<KIND, 1>(a, list, count, fail, s); break;
  case 1: launchWarp<KIND, 2>(a, list, count, fail, s); break;
  case 2: launchWarp<KIND, 3>(a, list, count, fail, s); break;
  case 3: launchWarp<KIND, 4>(a, list, count, fail, s); break;
  case 4: launchWarp<KIND, 5>(a, list, count, fail, s); break;
  case 5: launchWarp<KIND, 6>(a, list, count, fail, s); break;
  case 6: launchWarp<KIND, 7>(a, list, count, fail, s); break;
  default: launchWarp<KIND, 8>(a, list, count, fail, s); break;
  }
  RBF_LAUNCHED();
}

} // namespace

// buckets 0..7 (clusters of up to 8 (bucket + 1) <= 64 vertices)
void pumFactorWarp(const PumSolveArgs &a, int bucket, const PumEntry *list, int64_t count, int *fail, cudaStream_t s)
{
  RBF_DISPATCH_HOT_BASIS(a.rbf.kind, warpBucket<KIND>(a, bucket, list, count, fail, s));
}

} // namespace rbfmap
