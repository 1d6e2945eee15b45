// dense_lu.cu — blocked LU factorisation with partial pivoting and the matching solves, for the basis functions that are
// not positive definite (thin-plate splines, multiquadrics, volume splines) and for the integrated polynomial
// (polynomial="on", RadialBasisFctSolver.hpp:140-162, 172-205). Replaces Eigen::ColPivHouseholderQR as the
// decomposition of C (RadialBasisFctSolver.hpp:356-358, solve at :420-423, :452): both are backward-stable direct
// solvers of the same square system; singularity is reported like `!isInvertible()` (pivot below eps * N * max pivot).
// Panel factorisation in one CTA per 64 columns, row swaps, triangular panel solve, trailing update on the FP64 tensor
// cores through the same DMMA GEMM as the Cholesky path (dense.cu).
#include "dense.cuh"

#include <climits>

namespace rbfmap {

// from dense.cu
void denseGemmNtSubtract(const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int64_t rows, int64_t cols, int K, cudaStream_t s);

namespace {

constexpr int kLuNb = 64;

// factor the panel A(j0:np, j0:j0+64) in place with partial pivoting (LAPACK getf2 on a tall panel)
__global__ void __launch_bounds__(1024) luPanelKernel(double *__restrict__ A, long long ld, long long np, long long j0, int *__restrict__ piv,
                                                      double *__restrict__ pivStats /*[0] max |pivot|, [1] min |pivot|*/)
{
  __shared__ double redV[32];
  __shared__ int    redI[32];
  __shared__ double rowBuf[kLuNb]; // pivot row of the panel
  __shared__ int    pivRow;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int c = 0; c < kLuNb; ++c) {
    const long long col = j0 + c;
    // ---- pivot search: max |a(i, col)|, i >= col (ties -> lowest row) ----
    double best = -1.0;
    int    bi   = INT_MAX;
    for (long long i = col + tid; i < np; i += 1024) {
      const double v = fabs(A[i + col * ld]);
      if (v > best) {
        best = v;
        bi   = static_cast<int>(i);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int    oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi   = oi;
      }
    }
    if (lane == 0) {
      redV[w] = best;
      redI[w] = bi;
    }
    __syncthreads();
    if (w == 0) {
      best = redV[lane];
      bi   = redI[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int    oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) {
          best = ov;
          bi   = oi;
        }
      }
      if (lane == 0) {
        pivRow   = bi;
        piv[col] = bi;
        pivStats[0] = fmax(pivStats[0], best);
        pivStats[1] = fmin(pivStats[1], best);
      }
    }
    __syncthreads();
    const long long pr = pivRow;
    // ---- swap rows col <-> pr inside the panel, keep the pivot row in shared memory ----
    if (tid < kLuNb) {
      const double a = A[col + (j0 + tid) * ld], b = A[pr + (j0 + tid) * ld];
      A[col + (j0 + tid) * ld] = b;
      A[pr + (j0 + tid) * ld]  = a;
      rowBuf[tid]              = b;
    }
    __syncthreads();
    const double pv = rowBuf[c];
    if (pv != 0.0) {
      const double rp = 1.0 / pv;
      // ---- scale the column and update the remaining panel columns, one row per thread ----
      for (long long i = col + 1 + tid; i < np; i += 1024) {
        const double l  = A[i + col * ld] * rp;
        A[i + col * ld] = l;
        for (int cc = c + 1; cc < kLuNb; ++cc)
          A[i + (j0 + cc) * ld] = fma(-l, rowBuf[cc], A[i + (j0 + cc) * ld]);
      }
    }
    __syncthreads();
  }
}

// apply the 64 row interchanges of a panel to the columns outside it
__global__ void luSwapKernel(double *__restrict__ A, long long ld, long long np, long long j0, const int *__restrict__ piv)
{
  long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (j >= np - kLuNb)
    return;
  if (j >= j0)
    j += kLuNb; // skip the panel itself
  double *col = A + j * ld;
  for (int c = 0; c < kLuNb; ++c) {
    const long long r = j0 + c, p = piv[r];
    if (p != r) {
      const double t = col[r];
      col[r]         = col[p];
      col[p]         = t;
    }
  }
}

// U12 = L11^-1 A12 (unit lower 64 x 64), one column per thread; also writes U12^T into W for the NT GEMM
__global__ void __launch_bounds__(128) luTrsmKernel(double *__restrict__ A, long long ld, long long np, long long j0, double *__restrict__ W, long long ldw)
{
  __shared__ double L[kLuNb][kLuNb + 1];
  for (int e = threadIdx.x; e < kLuNb * kLuNb; e += blockDim.x) {
    const int i = e % kLuNb, k = e / kLuNb;
    L[i][k]     = A[(j0 + i) + (j0 + k) * ld];
  }
  __syncthreads();
  const long long jj = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; // column index relative to j0 + 64
  const long long j  = j0 + kLuNb + jj;
  if (j >= np)
    return;
  double  x[kLuNb];
  double *col = A + j0 + j * ld;
#pragma unroll
  for (int i = 0; i < kLuNb; ++i) {
    double s = col[i];
#pragma unroll
    for (int k = 0; k < i; ++k)
      s = fma(-L[i][k], x[k], s);
    x[i] = s;
  }
#pragma unroll
  for (int i = 0; i < kLuNb; ++i) {
    col[i]          = x[i];
    W[jj + i * ldw] = x[i];
  }
}

// ---- solves: right-hand sides planar B[c*ldb + i] ----
__global__ void luPermuteKernel(double *__restrict__ B, long long ldb, int nrhs, long long np, const int *__restrict__ piv)
{
  const int c = blockIdx.x;
  if (c >= nrhs || threadIdx.x != 0)
    return;
  double *b = B + c * ldb;
  for (long long r = 0; r < np; ++r) { // sequential interchanges (they do not commute)
    const long long p = piv[r];
    if (p != r) {
      const double t = b[r];
      b[r]           = b[p];
      b[p]           = t;
    }
  }
}

// substitution inside one 64-block (single CTA): forward = unit lower triangle, backward = upper triangle
template <bool FORWARD>
__global__ void __launch_bounds__(64) luSolveDiagKernel(const double *__restrict__ A, long long ld, long long k0, double *__restrict__ B, long long ldb, int nrhs)
{
  __shared__ double T[kLuNb][kLuNb + 1];
  __shared__ double xs[kLuNb];
  const int tid = threadIdx.x;
  for (int e = tid; e < kLuNb * kLuNb; e += 64) {
    const int i = e % kLuNb, k = e / kLuNb;
    T[i][k]     = A[(k0 + i) + (k0 + k) * ld];
  }
  for (int c = 0; c < nrhs; ++c) {
    double *b = B + c * ldb;
    __syncthreads();
    xs[tid] = b[k0 + tid];
    __syncthreads();
    if (FORWARD) {
      for (int k = 0; k < kLuNb; ++k) {
        const double xk = xs[k];
        if (tid > k)
          xs[tid] = fma(-T[tid][k], xk, xs[tid]);
        __syncthreads();
      }
    } else {
      for (int k = kLuNb - 1; k >= 0; --k) {
        if (tid == k)
          xs[k] = xs[k] / T[k][k];
        __syncthreads();
        const double xk = xs[k];
        if (tid < k)
          xs[tid] = fma(-T[tid][k], xk, xs[tid]);
        __syncthreads();
      }
    }
    b[k0 + tid] = xs[tid];
  }
}

// update of the other rows with the solved block: forward -> rows below, backward -> rows above
template <bool FORWARD>
__global__ void __launch_bounds__(256) luSolveUpdateKernel(const double *__restrict__ A, long long ld, long long np, long long k0, double *__restrict__ B,
                                                           long long ldb, int nrhs)
{
  __shared__ double xs[kLuNb];
  const int         tid = threadIdx.x;
  const long long   lo = FORWARD ? k0 + kLuNb : 0, hi = FORWARD ? np : k0;
  for (int c = 0; c < nrhs; ++c) {
    double *b = B + c * ldb;
    __syncthreads();
    if (tid < kLuNb)
      xs[tid] = b[k0 + tid];
    __syncthreads();
    for (long long i = lo + blockIdx.x * 256LL + tid; i < hi; i += static_cast<long long>(gridDim.x) * 256) {
      double acc = b[i];
#pragma unroll 8
      for (int k = 0; k < kLuNb; ++k)
        acc = fma(-A[i + (k0 + k) * ld], xs[k], acc);
      b[i] = acc;
    }
  }
}

} // namespace

// In-place LU of the np x np matrix (np multiple of 256). piv: np ints. Returns false if numerically singular.
bool denseLu(double *dA, int64_t np, int *dPiv, cudaStream_t s)
{
  DevBuf<double> W, st;
  W.alloc(static_cast<size_t>(np) * kLuNb);
  st.alloc(2);
  const double init[2] = {0.0, 1.7976931348623157e308};
  RBF_CUDA(cudaMemcpyAsync(st.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
  for (int64_t j0 = 0; j0 < np; j0 += kLuNb) {
    luPanelKernel<<<1, 1024, 0, s>>>(dA, np, np, j0, dPiv, st.p);
    RBF_LAUNCHED();
    luSwapKernel<<<divUp(np - kLuNb, 256), 256, 0, s>>>(dA, np, np, j0, dPiv);
    RBF_LAUNCHED();
    const int64_t rest = np - j0 - kLuNb;
    if (rest > 0) {
      luTrsmKernel<<<divUp(rest, 128), 128, 0, s>>>(dA, np, np, j0, W.p, np);
      RBF_LAUNCHED();
      // A22 -= L21 * U12 = L21 * (U12^T)^T
      denseGemmNtSubtract(dA + (j0 + kLuNb) + j0 * np, np, W.p, np, dA + (j0 + kLuNb) + (j0 + kLuNb) * np, np, rest, rest, kLuNb, s);
    }
  }
  RBF_CUDA(cudaGetLastError());
  double h[2];
  RBF_CUDA(cudaMemcpyAsync(h, st.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  // ColPivHouseholderQR::isInvertible(): every pivot above max pivot * eps * size (RadialBasisFctSolver.hpp:356-358)
  return h[1] > h[0] * 2.220446049250313e-16 * static_cast<double>(np);
}

void denseLuSolve(const double *dLU, int64_t np, const int *dPiv, double *dB, int64_t ldb, int nrhs, cudaStream_t s)
{
  luPermuteKernel<<<nrhs, 32, 0, s>>>(dB, ldb, nrhs, np, dPiv);
  RBF_LAUNCHED();
  for (int64_t k0 = 0; k0 < np; k0 += kLuNb) {
    luSolveDiagKernel<true><<<1, 64, 0, s>>>(dLU, np, k0, dB, ldb, nrhs);
    RBF_LAUNCHED();
    const int64_t rest = np - k0 - kLuNb;
    if (rest > 0) {
      luSolveUpdateKernel<true><<<divUp(rest, 256), 256, 0, s>>>(dLU, np, np, k0, dB, ldb, nrhs);
      RBF_LAUNCHED();
    }
  }
  for (int64_t k0 = np - kLuNb; k0 >= 0; k0 -= kLuNb) {
    luSolveDiagKernel<false><<<1, 64, 0, s>>>(dLU, np, k0, dB, ldb, nrhs);
    RBF_LAUNCHED();
    if (k0 > 0) {
      luSolveUpdateKernel<false><<<divUp(k0, 256), 256, 0, s>>>(dLU, np, np, k0, dB, ldb, nrhs);
      RBF_LAUNCHED();
    }
  }
  RBF_CUDA(cudaGetLastError());
}

} // namespace rbfmap
