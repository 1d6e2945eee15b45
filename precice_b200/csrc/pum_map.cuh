// pum_map.cuh — pieces shared by the consistent and the conservative per-cluster map kernels (sm_100a).
//
// Shared-memory tile of one cluster (doubles):
//   ms : packed factor, n(n+1)/2 rounded up to an even count; staged with one TMA bulk copy
//        (cp.async.bulk global -> shared, completion on an mbarrier) that overlaps the gather and the polynomial fit
//   vt : vertex tile, n rows of LDV doubles: x, y, z, then the NC right-hand sides / coefficients of the row
//   red: reduction scratch
#pragma once

#include "pum_device.cuh"

#include <algorithm>
#include <climits>

namespace rbfmap {

// row length of the vertex tile: (x, y, z, b0) or (x, y, z, b0, b1, b2) -> 2 or 3 aligned 16-byte loads per row
template <int NC>
struct TileLd {
  static constexpr int value = NC == 1 ? 4 : 6;
};

__host__ __device__ inline long long paddedTri(long long n) { return (n * (n + 1) / 2 + 1) & ~1LL; }

// ---- TMA bulk copy + mbarrier (PTX ISA: cp.async.bulk / mbarrier, sm_90+) ----
__device__ __forceinline__ unsigned smemAddr(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void     mbarInit(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulkLoad(void *dstSmem, const void *srcGlobal, unsigned bytes, unsigned long long *bar)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal),
               "r"(bytes), "r"(smemAddr(bar))
               : "memory");
}
// asks the memory system to bring `bytes` (multiple of 16, 16-byte aligned) from global memory into L2; no completion to wait for
__device__ __forceinline__ void bulkPrefetchL2(const void *srcGlobal, unsigned bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(srcGlobal), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long *bar, unsigned phase)
{
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smemAddr(bar)), "r"(phase)
                 : "memory");
  }
}

// In-place solve of (M D^-1 M^T) x = b for NC right-hand sides stored in the vertex tile (b(i, cc) = v[i*LDV + cc]).
// `mp` is the packed factor: column k = [1/d_k, M(k+1,k), ...]. Blocks of 8 columns: every thread redoes the tiny
// diagonal solve redundantly from broadcast reads, then updates its own rows — two barriers per block.
// IDX is int for the shared-memory copy (n <= 128) and long long for the global-memory variant.
template <int NC, int LDV, typename IDX>
__device__ __forceinline__ void blockedSolve(const double *__restrict__ mp, int n, double *v, int tid, int nThreads)
{
  // forward: M y = b ; b <- numerators u = D y
  for (int k0 = 0; k0 < n; k0 += 8) {
    const int nb = min(8, n - k0);
    IDX       base[8]; // M(i, k0+q) = mp[base[q] + i]
    base[0] = static_cast<IDX>(k0) * n - static_cast<IDX>(k0) * (k0 - 1) / 2 - k0;
#pragma unroll
    for (int q = 1; q < 8; ++q)
      base[q] = base[q - 1] + (n - (k0 + q - 1) - 1);
    double y[8][NC], un[8][NC];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < nb) {
        const int k = k0 + q;
        double    u[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          u[cc] = v[k * LDV + cc];
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2)
          if (q2 < q) {
            const double mkj = mp[base[q2] + k];
#pragma unroll
            for (int cc = 0; cc < NC; ++cc)
              u[cc] = fma(-mkj, y[q2][cc], u[cc]);
          }
        const double rd = mp[base[q] + k];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          y[q][cc]  = u[cc] * rd;
          un[q][cc] = u[cc];
        }
      }
    }
    __syncthreads(); // everyone has read b[k0..k0+nb) before the owners overwrite it with the numerators
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < nb && tid == (k0 + q) % nThreads) {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          v[(k0 + q) * LDV + cc] = un[q][cc];
      }
    for (int i = k0 + nb + tid; i < n; i += nThreads) {
      double acc[NC];
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        acc[cc] = v[i * LDV + cc];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < nb) {
          const double mik = mp[base[q] + i];
#pragma unroll
          for (int cc = 0; cc < NC; ++cc)
            acc[cc] = fma(-mik, y[q][cc], acc[cc]);
        }
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        v[i * LDV + cc] = acc[cc];
    }
    __syncthreads();
  }
  // backward: M^T x = u
  for (int k0 = ((n - 1) / 8) * 8; k0 >= 0; k0 -= 8) {
    const int nb = min(8, n - k0);
    IDX       base[8];
    base[0] = static_cast<IDX>(k0) * n - static_cast<IDX>(k0) * (k0 - 1) / 2 - k0;
#pragma unroll
    for (int q = 1; q < 8; ++q)
      base[q] = base[q - 1] + (n - (k0 + q - 1) - 1);
    double x[8][NC];
#pragma unroll
    for (int q = 7; q >= 0; --q) {
      if (q < nb) {
        const int k = k0 + q;
        double    u[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          u[cc] = v[k * LDV + cc];
#pragma unroll
        for (int q2 = 7; q2 >= 0; --q2)
          if (q2 > q && q2 < nb) {
            const double mik = mp[base[q] + k0 + q2]; // M(k0+q2, k)
#pragma unroll
            for (int cc = 0; cc < NC; ++cc)
              u[cc] = fma(-mik, x[q2][cc], u[cc]);
          }
        const double rd = mp[base[q] + k];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          x[q][cc] = u[cc] * rd;
      }
    }
    __syncthreads(); // everyone has read b[k0..k0+nb) before the owners overwrite it
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < nb && tid == (k0 + q) % nThreads) {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          v[(k0 + q) * LDV + cc] = x[q][cc];
      }
    for (int i = tid; i < k0; i += nThreads) {
      double    acc[NC];
      const IDX bi = static_cast<IDX>(i) * n - static_cast<IDX>(i) * (i - 1) / 2 - i; // M(k, i) = mp[bi + k]
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        acc[cc] = v[i * LDV + cc];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < nb) {
          const double mki = mp[bi + k0 + q];
#pragma unroll
          for (int cc = 0; cc < NC; ++cc)
            acc[cc] = fma(-mki, x[q][cc], acc[cc]);
        }
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        v[i * LDV + cc] = acc[cc];
    }
    __syncthreads();
  }
}

// Warp-level variant for factors staged in shared memory (n <= 32 * RPL): ONE warp performs both substitutions, lane l
// owns rows l, l + 32, ...; the solved entry of each step is broadcast with a shuffle, so no block barrier and no
// redundant diagonal solves are needed (about a quarter of the instructions of blockedSolve for n ~ 50).
// Forward: column access M(i, k) = mp[base_k + i] (contiguous over lanes); backward: row access M(k, i) = mp[base_i + k].
// Must be called by all 32 lanes of one warp; the other warps of the CTA wait at the caller's barrier.
template <int NC, int LDV, int RPL>
__device__ __forceinline__ void warpSolve(const double *__restrict__ mp, int n, double *v, int lane)
{
  double u[RPL][NC];
  int    bi[RPL]; // base of the column of each owned row: M(k, i) = mp[bi + k]
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int i = lane + 32 * r;
    bi[r]       = i * n - (i * (i - 1)) / 2 - i;
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      u[r][cc] = i < n ? v[i * LDV + cc] : 0.0;
  }
  // ---- forward: M y = b; afterwards u holds the numerators D y ----
  int base = 0; // base_k = colStart(k) - k
#pragma unroll
  for (int r0 = 0; r0 < RPL; ++r0) {
    const int kEnd = min(32, n - 32 * r0);
    for (int kk = 0; kk < kEnd; ++kk) {
      const int    k  = 32 * r0 + kk;
      const double rd = mp[base + k];
      double       y[NC];
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        y[cc] = __shfl_sync(0xffffffffu, u[r0][cc], kk) * rd;
#pragma unroll
      for (int r = r0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        if (i > k && i < n) {
          const double mik = mp[base + i];
#pragma unroll
          for (int cc = 0; cc < NC; ++cc)
            u[r][cc] = fma(-mik, y[cc], u[r][cc]);
        }
      }
      base += n - k - 1;
    }
  }
  // ---- backward: M^T x = u ----
#pragma unroll
  for (int r0 = RPL - 1; r0 >= 0; --r0) {
    const int kEnd = min(32, n - 32 * r0);
    for (int kk = kEnd - 1; kk >= 0; --kk) {
      const int k = 32 * r0 + kk;
      base -= n - k - 1; // back to base_k
      const double rd = mp[base + k];
      double       x[NC];
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        x[cc] = __shfl_sync(0xffffffffu, u[r0][cc], kk) * rd;
      if (lane == kk) {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          u[r0][cc] = x[cc]; // the owner keeps the solution entry
      }
#pragma unroll
      for (int r = 0; r <= r0; ++r) {
        const int i = lane + 32 * r;
        if (i < k) {
          const double mki = mp[bi[r] + k];
#pragma unroll
          for (int cc = 0; cc < NC; ++cc)
            u[r][cc] = fma(-mki, x[cc], u[r][cc]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int i = lane + 32 * r;
    if (i < n) {
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        v[i * LDV + cc] = u[r][cc];
    }
  }
}

// lambda = C^-1 f with the explicit inverse of pum_inverse.cu (S[j n + i] = (C^-1)(i, j), in shared or global memory):
// dense matrix-vector product in place of the two triangular solves. All threads of the CTA; v rows as in warpSolve.
template <int NC, int LDV>
__device__ __forceinline__ void inverseApply(const double *__restrict__ S, int n, double *v, int tid, int nThreads)
{
  constexpr int kRows = 4; // rows per thread and pass (n <= 4 * nThreads is guaranteed by the callers: n <= 128, >= 32 threads)
  double        acc[kRows][NC];
#pragma unroll
  for (int r = 0; r < kRows; ++r)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      acc[r][cc] = 0.0;
  for (int j = 0; j < n; ++j) {
    double f[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      f[cc] = v[j * LDV + cc];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int i = tid + r * nThreads;
      if (i < n) {
        const double s = S[static_cast<long long>(j) * n + i];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          acc[r][cc] = fma(s, f[cc], acc[r][cc]);
      }
    }
  }
  __syncthreads(); // every thread has read all of f
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int i = tid + r * nThreads;
    if (i < n) {
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        v[i * LDV + cc] = acc[r][cc];
    }
  }
  __syncthreads();
}

template <typename K>
inline void allowSmem(K kernel, size_t bytes)
{
  if (bytes > 48 * 1024)
    RBF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
}

} // namespace rbfmap
