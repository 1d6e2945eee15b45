// pum_device.cuh — device helpers shared by the per-cluster kernels (pum_poly.cu, pum_factor.cu, pum_map_*.cu).
#pragma once

#include "pum.cuh"

namespace rbfmap {

// packed lower-triangular column-major storage: start of column k of an n x n factor
__device__ __forceinline__ int colStart32(int k, int n) { return k * n - (k * (k - 1)) / 2; }
__device__ __forceinline__ long long colStart(int k, int n) { return static_cast<long long>(k) * n - static_cast<long long>(k) * (k - 1) / 2; }

// Basis function with the kind fixed at compile time (hot kinds) or dispatched at run time (KIND_RUNTIME).
constexpr int KIND_RUNTIME = -1;
template <int KIND>
__device__ __forceinline__ double evalBasisK(double r, const RbfParams &f)
{
  if (KIND == KIND_RUNTIME)
    return evalBasis(r, f);
  else
    return evalBasisT<KIND>(r, f);
}

// Host dispatch: compact polynomials get their own instantiation, every other function shares the run-time one.
#define RBF_DISPATCH_HOT_BASIS(kind, ...)                                                                         \
  switch (kind) {                                                                                                 \
  case RBFMAP_COMPACT_POLYNOMIAL_C0: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C0; __VA_ARGS__; } break;   \
  case RBFMAP_COMPACT_POLYNOMIAL_C2: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C2; __VA_ARGS__; } break;   \
  case RBFMAP_COMPACT_POLYNOMIAL_C4: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C4; __VA_ARGS__; } break;   \
  case RBFMAP_COMPACT_POLYNOMIAL_C6: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C6; __VA_ARGS__; } break;   \
  case RBFMAP_COMPACT_POLYNOMIAL_C8: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C8; __VA_ARGS__; } break;   \
  default: { constexpr int KIND = KIND_RUNTIME; __VA_ARGS__; } break;                                             \
  }

// Square root for squared vertex distances: MUFU.RSQ64H seed, one Newton step on the reciprocal root and one
// residual correction of the root. No special-case branches; the argument is clamped to >= 1e-300 so that x == 0
// (coincident vertices, the matrix diagonal) yields 1e-150, which every basis function maps to phi(0) exactly.
// Faithful (<= 1 ulp), and equal to the IEEE square root except for rare half-ulp ties (checked by
// tests/test_gpu_parity.py::test_device_sqrt).
__device__ __forceinline__ double distSqrt(double x)
{
  x = fmax(x, 1.0e-300);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  const double e  = fma(-hx * y, y, 0.5);
  y               = fma(y, e, y);           // ~2^-43 relative
  double       s  = x * y;                  // sqrt estimate
  const double r  = fma(-s, s, x);          // residual
  s               = fma(r, 0.5 * y, s);     // corrected root
  return s;
}

// squared distance, all axes (3-D) or x/y only (2-D meshes carry z = 0, activeAxis = {1,1,0});
// same operation order as RadialBasisFctSolver.hpp:102-113, no contraction
template <bool DIM3>
__device__ __forceinline__ double sqDistD(double ax, double ay, double az, double bx, double by, double bz)
{
  const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
  double       s  = __dmul_rn(dx, dx);
  s               = __dadd_rn(s, __dmul_rn(dy, dy));
  if (DIM3) {
    const double dz = __dsub_rn(az, bz);
    s               = __dadd_rn(s, __dmul_rn(dz, dz));
  }
  return s;
}

// ---- fast pair evaluation for the algebra kernels ----
// phi(|a - b|) of the compactly supported polynomial functions with fused multiply-adds: 20 FP64 instructions for C6
// in 3-D instead of the 33 of the operation-by-operation sequence above (the FP64 pipe is what bounds the per-cluster
// kernels). Differs from the reference sequence by a few ulp of phi(0); the parity tests bound the effect on mapped values.
//   x = |a-b|^2 / R^2 + 1e-300   (the offset keeps the reciprocal root finite on the diagonal)
//   x clamped to <= 1 with an integer compare on the high word: beyond the support (1-p)^k <= 1e-31 instead of 0
//   p = sqrt(x) = s (1 - e)^(-1/2), s = x y, e = 1 - s y, y = MUFU.RSQ64H seed (|e| < 2^-20): s (1 + e/2 + 3/8 e^2)
template <int KIND>
struct FastBasis {
  static constexpr bool value = KIND == RBFMAP_COMPACT_POLYNOMIAL_C0 || KIND == RBFMAP_COMPACT_POLYNOMIAL_C2 || KIND == RBFMAP_COMPACT_POLYNOMIAL_C4 ||
                                KIND == RBFMAP_COMPACT_POLYNOMIAL_C6 || KIND == RBFMAP_COMPACT_POLYNOMIAL_C8;
};

template <int KIND>
__device__ __forceinline__ double compactPolynomialOfSquare(double x)
{
  if (__double2hiint(x) >= 0x3ff00000) // x >= 1 (also inf / nan): integer compare and selects, off the FP64 pipe
    x = 1.0;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double s = x * y;
  const double e = fma(-s, y, 1.0);
  const double p = fma(s, e * fma(0.375, e, 0.5), s);
  const double q = 1.0 - p;
  const double q2 = q * q;
  if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C0) {
    return q2;
  } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C2) {
    return (q2 * q2) * fma(4.0, p, 1.0);
  } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C4) {
    return (q2 * q2 * q2) * fma(fma(35.0, p, 18.0), p, 3.0);
  } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C6) {
    const double q4 = q2 * q2;
    return (q4 * q4) * fma(fma(fma(32.0, p, 25.0), p, 8.0), p, 1.0);
  } else {
    const double q4 = q2 * q2;
    return (q4 * q4 * q2) * fma(fma(fma(fma(1287.0, p, 1350.0), p, 630.0), p, 150.0), p, 15.0);
  }
}

// invR2 = 1 / support^2 for the fast kinds (unused otherwise)
template <int KIND, bool DIM3>
__device__ __forceinline__ double pairBasis(double ax, double ay, double az, double bx, double by, double bz, const RbfParams &f, double invR2)
{
  if (FastBasis<KIND>::value) {
    const double dx = ax - bx, dy = ay - by;
    double       d2 = fma(dy, dy, dx * dx);
    if (DIM3) {
      const double dz = az - bz;
      d2              = fma(dz, dz, d2);
    }
    return compactPolynomialOfSquare<KIND>(fma(d2, invR2, 1.0e-300));
  } else {
    return evalBasisK<KIND>(distSqrt(sqDistD<DIM3>(ax, ay, az, bx, by, bz)), f);
  }
}

// N independent pair evaluations written stage by stage. A warp issues in order and a dependent FP64 operation waits
// ~16-20 cycles for its operand, so a single evaluation (a chain of ~20 dependent operations) uses a twentieth of the
// pipe; ptxas does not interleave unrolled iterations by itself, but it keeps this order. All N pairs share the point a.
template <int KIND, bool DIM3, int N>
__device__ __forceinline__ void pairBasisBatch(const double (&ax)[N], const double (&ay)[N], const double (&az)[N], const double (&bx)[N],
                                               const double (&by)[N], const double (&bz)[N], const RbfParams &f, double invR2, double (&out)[N])
{
  if (FastBasis<KIND>::value) {
    double x[N], y[N], s[N], e[N], p[N], q[N], q2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double dx = ax[i] - bx[i];
      x[i]            = dx * dx;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double dy = ay[i] - by[i];
      x[i]            = fma(dy, dy, x[i]);
    }
    if (DIM3) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double dz = az[i] - bz[i];
        x[i]            = fma(dz, dz, x[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
      x[i] = fma(x[i], invR2, 1.0e-300);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (__double2hiint(x[i]) >= 0x3ff00000)
        x[i] = 1.0;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(x[i]));
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
      s[i] = x[i] * y[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
      e[i] = fma(-s[i], y[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i)
      y[i] = fma(0.375, e[i], 0.5);
#pragma unroll
    for (int i = 0; i < N; ++i)
      e[i] = e[i] * y[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
      p[i] = fma(s[i], e[i], s[i]);
#pragma unroll
    for (int i = 0; i < N; ++i)
      q[i] = 1.0 - p[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
      q2[i] = q[i] * q[i];
    if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C0) {
#pragma unroll
      for (int i = 0; i < N; ++i)
        out[i] = q2[i];
    } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C2) {
#pragma unroll
      for (int i = 0; i < N; ++i)
        out[i] = (q2[i] * q2[i]) * fma(4.0, p[i], 1.0);
    } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C4) {
#pragma unroll
      for (int i = 0; i < N; ++i)
        out[i] = (q2[i] * q2[i] * q2[i]) * fma(fma(35.0, p[i], 18.0), p[i], 3.0);
    } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C6) {
      double h[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        q2[i] = q2[i] * q2[i]; // q^4
        h[i]  = fma(32.0, p[i], 25.0);
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        q2[i] = q2[i] * q2[i]; // q^8
        h[i]  = fma(h[i], p[i], 8.0);
      }
#pragma unroll
      for (int i = 0; i < N; ++i)
        h[i] = fma(h[i], p[i], 1.0);
#pragma unroll
      for (int i = 0; i < N; ++i)
        out[i] = q2[i] * h[i];
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double q4 = q2[i] * q2[i];
        out[i]          = (q4 * q4 * q2[i]) * fma(fma(fma(fma(1287.0, p[i], 1350.0), p[i], 630.0), p[i], 150.0), p[i], 15.0);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i)
      out[i] = evalBasisK<KIND>(distSqrt(sqDistD<DIM3>(ax[i], ay[i], az[i], bx[i], by[i], bz[i])), f);
  }
}

// all N pairs share the point a
template <int KIND, bool DIM3, int N>
__device__ __forceinline__ void pairBasisBatch(double ax, double ay, double az, const double (&bx)[N], const double (&by)[N], const double (&bz)[N],
                                               const RbfParams &f, double invR2, double (&out)[N])
{
  double axs[N], ays[N], azs[N];
#pragma unroll
  for (int i = 0; i < N; ++i)
    axs[i] = ax, ays[i] = ay, azs[i] = az;
  pairBasisBatch<KIND, DIM3, N>(axs, ays, azs, bx, by, bz, f, invR2, out);
}

// reciprocal of a positive pivot: MUFU.RCP64H seed + two Newton steps, no slow path (~1 ulp)
__device__ __forceinline__ double pivotReciprocal(double d)
{
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, 1.0);
  y        = fma(y, e, y);
  e        = fma(-d, y, 1.0);
  return fma(y, e, y);
}

// polynomial basis row of a vertex for a cluster record (mode 1: centred/scaled, mode 2: raw)
__device__ __forceinline__ void polyRow(int mode, int mask, double x, double y, double z, double cx, double cy, double cz, double rinv, double (&q)[4])
{
  q[0] = 1.0;
  q[1] = q[2] = q[3] = 0.0;
  double v[3];
  if (mode == 1) {
    v[0] = (x - cx) * rinv;
    v[1] = (y - cy) * rinv;
    v[2] = (z - cz) * rinv;
  } else {
    v[0] = x;
    v[1] = y;
    v[2] = z;
  }
  int k = 1;
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if (mask & (1 << d)) {
      if (k == 1) q[1] = v[d];
      else if (k == 2) q[2] = v[d];
      else q[3] = v[d];
      ++k;
    }
}

// block-wide sum of NV values per thread; result broadcast to all threads through `red` (>= NV * nWarps doubles)
template <int NV>
__device__ __forceinline__ void blockReduceSum(double (&v)[NV], double *red, int nThreads)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nThreads >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
  __syncthreads(); // `red` may still be read from a previous reduction
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q)
      red[w * NV + q] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double s = 0.0;
    for (int ww = 0; ww < nw; ++ww)
      s += red[ww * NV + q];
    v[q] = s;
  }
}

// normalized partition-of-unity weight of out-vertex (x,y,z) in the cluster centred at (cx,cy,cz)
// (PartitionOfUnityMapping.hpp:321-336)
// `zeroCount` points to the number of covering clusters whose weight is exactly zero; it is only read when the weight sum
// is not positive - then every covering cluster has weight zero and the count is the number of covering clusters
__device__ __forceinline__ double normalizedWeight(double x, double y, double z, double cx, double cy, double cz, double rinv, double wsum,
                                                   const int *__restrict__ zeroCount)
{
  if (wsum <= 0.0)
    return __ddiv_rn(1.0, static_cast<double>(*zeroCount));
  const double w = pumWeight(__dsqrt_rn(sqDist3(cx, cy, cz, x, y, z)), rinv);
  return __ddiv_rn(w, wsum);
}

// coordinate of padding vertex i (beyond the cluster size): far away and pairwise distinct, so that every basis
// function evaluates to (numerically) zero against real vertices without any per-entry conditional
__device__ __forceinline__ double paddingCoordinate(int i) { return 1.0e100 * static_cast<double>(i + 1); }

} // namespace rbfmap
