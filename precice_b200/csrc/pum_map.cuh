// pum_map.cuh — pieces shared by the consistent and the conservative map kernels.
#pragma once

#include "pum_device.cuh"

#include <algorithm>
#include <climits>

namespace rbfmap {

// In-place solve of (M D^-1 M^T) x = b for NC right-hand sides held in shared memory (b[comp*ldb + i]).
// `mp` is the packed factor (shared or global memory). Blocks of 8 columns: every thread redoes the tiny
// diagonal solve redundantly (broadcast reads), then updates its own row — one barrier per block.
template <int NC>
__device__ __forceinline__ void blockedSolve(const double *mp, int n, double *b, int ldb, int tid, int nThreads)
{
  // forward: M y = b ; b <- numerators u = D y
  for (int k0 = 0; k0 < n; k0 += 8) {
    const int nb = min(8, n - k0);
    double    y[8][NC], un[8][NC];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < nb) {
        const int k = k0 + q;
        double    u[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          u[cc] = b[cc * ldb + k];
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2)
          if (q2 < q) {
            const double mkj = mp[colStart(k0 + q2, n) + (k - k0 - q2)];
#pragma unroll
            for (int cc = 0; cc < NC; ++cc)
              u[cc] = fma(-mkj, y[q2][cc], u[cc]);
          }
        const double rd = mp[colStart(k, n)];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          y[q][cc] = u[cc] * rd;
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          un[q][cc] = u[cc];
      }
    }
    __syncthreads(); // everyone has read b[k0..k0+nb) before the owners overwrite it with the numerators
    for (int q = 0; q < nb; ++q)
      if (tid == (k0 + q) % nThreads) {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          b[cc * ldb + k0 + q] = un[q][cc];
      }
    for (int i = k0 + nb + tid; i < n; i += nThreads) {
      double acc[NC];
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        acc[cc] = b[cc * ldb + i];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < nb) {
          const double mik = mp[colStart(k0 + q, n) + (i - k0 - q)];
#pragma unroll
          for (int cc = 0; cc < NC; ++cc)
            acc[cc] = fma(-mik, y[q][cc], acc[cc]);
        }
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        b[cc * ldb + i] = acc[cc];
    }
    __syncthreads();
  }
  // backward: M^T x = u
  for (int k0 = ((n - 1) / 8) * 8; k0 >= 0; k0 -= 8) {
    const int nb = min(8, n - k0);
    double    x[8][NC];
#pragma unroll
    for (int q = 7; q >= 0; --q) {
      if (q < nb) {
        const int k = k0 + q;
        double    u[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          u[cc] = b[cc * ldb + k];
#pragma unroll
        for (int q2 = 7; q2 >= 0; --q2)
          if (q2 > q && q2 < nb) {
            const double mik = mp[colStart(k, n) + (q2 - q)]; // M(k0+q2, k)
#pragma unroll
            for (int cc = 0; cc < NC; ++cc)
              u[cc] = fma(-mik, x[q2][cc], u[cc]);
          }
        const double rd = mp[colStart(k, n)];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          x[q][cc] = u[cc] * rd;
      }
    }
    __syncthreads(); // everyone has read b[k0..k0+nb) before the owners overwrite it
    for (int q = 0; q < nb; ++q)
      if (tid == (k0 + q) % nThreads) {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          b[cc * ldb + k0 + q] = x[q][cc];
      }
    for (int i = tid; i < k0; i += nThreads) {
      double          acc[NC];
      const long long ci = colStart(i, n);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        acc[cc] = b[cc * ldb + i];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < nb) {
          const double mki = mp[ci + (k0 + q - i)]; // M(k0+q, i)
#pragma unroll
          for (int cc = 0; cc < NC; ++cc)
            acc[cc] = fma(-mki, x[q][cc], acc[cc]);
        }
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        b[cc * ldb + i] = acc[cc];
    }
    __syncthreads();
  }
}


template <typename K>
inline void allowSmem(K kernel, size_t bytes)
{
  if (bytes > 48 * 1024)
    RBF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
}

// threads per CTA of the map kernels for a bucket, and whether the packed factor is staged in shared memory
inline int mapThreads(int bucket) { return bucket < 8 ? 64 : (bucket < 12 ? 128 : 256); }

} // namespace rbfmap
