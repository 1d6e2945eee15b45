// basis.cuh — the eleven radial basis functions of the hot path as device functors.
//
// Replaces mapping/impl/BasisFunctions.hpp:87-607 (reference functors) on the device. Each
// evaluation reproduces the reference's floating-point operation sequence (pow_int by repeated
// squaring, math/math.hpp:21-40; explicit fma only where the reference has PRECICE_FMA) using
// non-contracting intrinsics, so that the polynomial functions are BIT-IDENTICAL to the CPU
// path; exp/log based functions agree to the accuracy of the CUDA math library (<= 1 ulp).
#pragma once

#include "common.cuh"

#include <cmath>
#include <limits>

namespace rbfmap {

// impl/BasisFunctions.hpp:40-44 (RadialBasisParameters) + the function selector
struct RbfParams {
  int    kind;
  double p1, p2, p3;
};

inline bool basisIsSpd(int kind)
{
  return !(kind == RBFMAP_THIN_PLATE_SPLINES || kind == RBFMAP_MULTIQUADRICS || kind == RBFMAP_VOLUME_SPLINES);
}
inline bool basisHasCompactSupport(int kind)
{
  return !(kind == RBFMAP_THIN_PLATE_SPLINES || kind == RBFMAP_MULTIQUADRICS || kind == RBFMAP_INVERSE_MULTIQUADRICS || kind == RBFMAP_VOLUME_SPLINES);
}

// Host-side construction: same parameter derivation and the same argument checks / messages as the
// reference functor constructors (BasisFunctions.hpp:120-124, 159-166, 229-252, 307-316, ...).
RbfParams makeRbfParams(int kind, double supportRadius, double shapeParameter);
double    basisSupportRadius(const RbfParams &f);

#ifdef __CUDACC__
namespace dev {
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double pow2(double x) { return mul(x, x); }
__device__ __forceinline__ double pow3(double x) { return mul(mul(x, x), x); }
__device__ __forceinline__ double pow4(double x)
{
  const double x2 = mul(x, x);
  return mul(x2, x2);
}
__device__ __forceinline__ double pow5(double x)
{
  const double x2 = mul(x, x);
  return mul(mul(x2, x2), x);
}
__device__ __forceinline__ double pow6(double x)
{
  const double x2 = mul(x, x);
  return mul(mul(x2, x2), x2);
}
__device__ __forceinline__ double pow8(double x)
{
  const double x2 = mul(x, x), x4 = mul(x2, x2);
  return mul(x4, x4);
}
__device__ __forceinline__ double pow10(double x)
{
  const double x2 = mul(x, x), x4 = mul(x2, x2);
  return mul(mul(x4, x4), x2);
}
} // namespace dev

template <int KIND>
__device__ __forceinline__ double evalBasisT(double radius, const RbfParams &f)
{
  using namespace dev;
  if (KIND == RBFMAP_THIN_PLATE_SPLINES) { // :95-99
    return mul(log(fmax(radius, 1.0e-14)), pow2(radius));
  } else if (KIND == RBFMAP_MULTIQUADRICS) { // :131-135
    return sqrt(add(f.p1, pow2(radius)));
  } else if (KIND == RBFMAP_INVERSE_MULTIQUADRICS) { // :172-176
    return __ddiv_rn(1.0, sqrt(add(f.p1, pow2(radius))));
  } else if (KIND == RBFMAP_VOLUME_SPLINES) { // :205-208
    return fabs(radius);
  } else if (KIND == RBFMAP_GAUSSIAN) { // :263-272
    if (radius > f.p2)
      return 0.0;
    return sub(exp(-pow2(mul(f.p1, radius))), f.p3);
  } else {
    const double p = mul(radius, f.p1);
    if (p >= 1.0)
      return 0.0;
    if (KIND == RBFMAP_COMPACT_TPS_C2) { // :327-335
      double v = sub(1.0, mul(30.0, pow2(p)));
      v        = sub(v, mul(10.0, pow3(p)));
      v        = add(v, mul(45.0, pow4(p)));
      v        = sub(v, mul(6.0, pow5(p)));
      v        = sub(v, mul(mul(pow3(p), 60.0), log(fmax(p, 1.0e-14))));
      return v;
    } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C0) { // :381-388
      return pow2(sub(1.0, p));
    } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C2) { // :434-441
      return mul(pow4(sub(1.0, p)), fma(4.0, p, 1.0));
    } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C4) { // :487-494
      return mul(pow6(sub(1.0, p)), add(mul(35.0, pow2(p)), fma(18.0, p, 3.0)));
    } else if (KIND == RBFMAP_COMPACT_POLYNOMIAL_C6) { // :538-545
      return mul(pow8(sub(1.0, p)), add(add(mul(32.0, pow3(p)), mul(25.0, pow2(p))), fma(8.0, p, 1.0)));
    } else { // RBFMAP_COMPACT_POLYNOMIAL_C8  :591-598
      double q = add(mul(1287.0, pow4(p)), mul(1350.0, pow3(p)));
      q        = add(q, mul(630.0, pow2(p)));
      q        = add(q, mul(150.0, p));
      q        = add(q, 15.0);
      return mul(pow10(sub(1.0, p)), q);
    }
  }
}

// runtime dispatch (cold paths)
__device__ __forceinline__ double evalBasis(double radius, const RbfParams &f)
{
  switch (f.kind) {
  case RBFMAP_COMPACT_POLYNOMIAL_C0: return evalBasisT<RBFMAP_COMPACT_POLYNOMIAL_C0>(radius, f);
  case RBFMAP_COMPACT_POLYNOMIAL_C2: return evalBasisT<RBFMAP_COMPACT_POLYNOMIAL_C2>(radius, f);
  case RBFMAP_COMPACT_POLYNOMIAL_C4: return evalBasisT<RBFMAP_COMPACT_POLYNOMIAL_C4>(radius, f);
  case RBFMAP_COMPACT_POLYNOMIAL_C6: return evalBasisT<RBFMAP_COMPACT_POLYNOMIAL_C6>(radius, f);
  case RBFMAP_COMPACT_POLYNOMIAL_C8: return evalBasisT<RBFMAP_COMPACT_POLYNOMIAL_C8>(radius, f);
  case RBFMAP_THIN_PLATE_SPLINES: return evalBasisT<RBFMAP_THIN_PLATE_SPLINES>(radius, f);
  case RBFMAP_MULTIQUADRICS: return evalBasisT<RBFMAP_MULTIQUADRICS>(radius, f);
  case RBFMAP_INVERSE_MULTIQUADRICS: return evalBasisT<RBFMAP_INVERSE_MULTIQUADRICS>(radius, f);
  case RBFMAP_VOLUME_SPLINES: return evalBasisT<RBFMAP_VOLUME_SPLINES>(radius, f);
  case RBFMAP_GAUSSIAN: return evalBasisT<RBFMAP_GAUSSIAN>(radius, f);
  default: return evalBasisT<RBFMAP_COMPACT_TPS_C2>(radius, f);
  }
}

// Wendland C2 partition-of-unity weight, support = cluster radius (SphericalVertexCluster.hpp:146, 337-344)
__device__ __forceinline__ double pumWeight(double dist, double rinv)
{
  RbfParams w{RBFMAP_COMPACT_POLYNOMIAL_C2, rinv, 0.0, 0.0};
  return evalBasisT<RBFMAP_COMPACT_POLYNOMIAL_C2>(dist, w);
}
#endif

// Host dispatch helper: calls fn.template operator()<KIND>() for the runtime kind.
#define RBF_DISPATCH_BASIS(kind, ...)                                                  \
  switch (kind) {                                                                      \
  case RBFMAP_COMPACT_POLYNOMIAL_C0: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C0; __VA_ARGS__; } break; \
  case RBFMAP_COMPACT_POLYNOMIAL_C2: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C2; __VA_ARGS__; } break; \
  case RBFMAP_COMPACT_POLYNOMIAL_C4: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C4; __VA_ARGS__; } break; \
  case RBFMAP_COMPACT_POLYNOMIAL_C6: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C6; __VA_ARGS__; } break; \
  case RBFMAP_COMPACT_POLYNOMIAL_C8: { constexpr int KIND = RBFMAP_COMPACT_POLYNOMIAL_C8; __VA_ARGS__; } break; \
  case RBFMAP_THIN_PLATE_SPLINES: { constexpr int KIND = RBFMAP_THIN_PLATE_SPLINES; __VA_ARGS__; } break;       \
  case RBFMAP_MULTIQUADRICS: { constexpr int KIND = RBFMAP_MULTIQUADRICS; __VA_ARGS__; } break;                 \
  case RBFMAP_INVERSE_MULTIQUADRICS: { constexpr int KIND = RBFMAP_INVERSE_MULTIQUADRICS; __VA_ARGS__; } break; \
  case RBFMAP_VOLUME_SPLINES: { constexpr int KIND = RBFMAP_VOLUME_SPLINES; __VA_ARGS__; } break;               \
  case RBFMAP_GAUSSIAN: { constexpr int KIND = RBFMAP_GAUSSIAN; __VA_ARGS__; } break;                           \
  default: { constexpr int KIND = RBFMAP_COMPACT_TPS_C2; __VA_ARGS__; } break;                                  \
  }

} // namespace rbfmap
