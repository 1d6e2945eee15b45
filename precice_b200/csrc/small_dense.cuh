// small_dense.cuh — tiny dense helpers shared by the device kernels and the host-side global mappings.
#pragma once

#include <cmath>

// the helpers below are also compiled for the host, whose compiler does not know the unroll pragma
#ifdef __CUDA_ARCH__
#define RBF_UNROLL _Pragma("unroll")
#define RBF_NO_UNROLL _Pragma("unroll 1")
#else
#define RBF_UNROLL
#define RBF_NO_UNROLL
#endif

namespace rbfmap {

// eigenvalues of a symmetric 4x4 matrix (cyclic Jacobi, everything in registers)
__host__ __device__ inline void jacobiEigenvalues4(double a[4][4], double &evMin, double &evMax)
{
RBF_NO_UNROLL
  for (int sweep = 0; sweep < 16; ++sweep) {
    double off = 0.0;
RBF_UNROLL
    for (int p = 0; p < 3; ++p)
RBF_UNROLL
      for (int q = p + 1; q < 4; ++q)
        off += a[p][q] * a[p][q];
    if (off == 0.0)
      break;
RBF_UNROLL
    for (int p = 0; p < 3; ++p)
RBF_UNROLL
      for (int q = p + 1; q < 4; ++q) {
        const double apq = a[p][q];
        if (fabs(apq) > 1e-300) {
          const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
          const double t     = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c     = 1.0 / sqrt(t * t + 1.0);
          const double s     = t * c;
          a[p][p] -= t * apq;
          a[q][q] += t * apq;
          a[p][q] = 0.0;
          a[q][p] = 0.0;
RBF_UNROLL
          for (int r = 0; r < 4; ++r)
            if (r != p && r != q) {
              const double arp = a[r][p], arq = a[r][q];
              a[r][p] = a[p][r] = c * arp - s * arq;
              a[r][q] = a[q][r] = s * arp + c * arq;
            }
        }
      }
  }
  evMin = fmin(fmin(a[0][0], a[1][1]), fmin(a[2][2], a[3][3]));
  evMax = fmax(fmax(a[0][0], a[1][1]), fmax(a[2][2], a[3][3]));
}


// inverse of a symmetric positive definite 4x4 matrix by Gauss-Jordan elimination (no pivoting needed); returns
// ||G||_inf * ||G^-1||_inf as a condition estimate
__host__ __device__ inline double invertSpd4(double g[4][4], double inv[4][4])
{
  double normG = 0.0;
  for (int r = 0; r < 4; ++r)
    normG = fmax(normG, fabs(g[r][0]) + fabs(g[r][1]) + fabs(g[r][2]) + fabs(g[r][3]));
  for (int r = 0; r < 4; ++r)
    for (int s = 0; s < 4; ++s)
      inv[r][s] = r == s ? 1.0 : 0.0;
  for (int k = 0; k < 4; ++k) {
    const double piv = 1.0 / g[k][k];
    for (int s = 0; s < 4; ++s) {
      g[k][s] *= piv;
      inv[k][s] *= piv;
    }
    for (int r = 0; r < 4; ++r)
      if (r != k) {
        const double f = g[r][k];
        for (int s = 0; s < 4; ++s) {
          g[r][s] -= f * g[k][s];
          inv[r][s] -= f * inv[k][s];
        }
      }
  }
  double normI = 0.0;
  for (int r = 0; r < 4; ++r)
    normI = fmax(normI, fabs(inv[r][0]) + fabs(inv[r][1]) + fabs(inv[r][2]) + fabs(inv[r][3]));
  return normG * normI;
}

} // namespace rbfmap
