// pum_inverse.cu — rbf-pum-direct with basis functions that are not strictly positive definite (thin-plate splines,
// multiquadrics, volume splines): per-cluster assembly of the full interpolation matrix and its explicit inverse (sm_100a).
//
// Replaces, per cluster, RadialBasisFctSolver.hpp:164-207 (buildMatrixCLU) and the non-SPD branch of the constructor,
// :354-358 (Eigen::ColPivHouseholderQR of C) with the invertibility check of :356-365; cluster constructor
// SphericalVertexCluster.hpp:139-182. The reference's own device path rejects these functions (BatchedRBFSolver.hpp:113).
//
// C is symmetric but indefinite here (its diagonal is phi(0) = 0 for thin-plate and volume splines), so the square-root-
// free Cholesky kernels do not apply. One CTA per cluster: C is assembled in shared memory and inverted in place by
// Gauss-Jordan elimination with partial (row) pivoting; the map kernels then apply C^-1 as a dense matrix-vector product
// instead of two triangular solves. Stored per cluster: n x n doubles, S[j * n + i] = (C^-1)(i, j).
// A pivot below eps * n * (largest pivot) counts as singular (the rank criterion of ColPivHouseholderQR::isInvertible
// with Eigen's default threshold) and raises the reference's "not invertable" error.
#include "pum_device.cuh"

#include <climits>

namespace rbfmap {
namespace {

constexpr int kInvThreads = 128;

template <bool DIM3>
__global__ void __launch_bounds__(kInvThreads) pumInverseKernel(PumSolveArgs args, const PumEntry *__restrict__ list, int *__restrict__ fail)
{
  extern __shared__ __align__(16) double dyn[]; // a: n x n (column-major), rowk: n, colk: n, sx, sy, sz: 3 n
  __shared__ double redVal[kInvThreads / 32];
  __shared__ int    redIdx[kInvThreads / 32];
  __shared__ int    pivRow;
  __shared__ double pivMax;
  __shared__ int    bad;
  const PumEntry  en  = list[blockIdx.x];
  const int       n   = en.n;
  const int       tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const RbfParams rbf = args.rbf;
  double         *a = dyn, *rowk = a + static_cast<size_t>(n) * n, *colk = rowk + n, *sx = colk + n, *sy = sx + n, *sz = sy + n;
  int            *perm = reinterpret_cast<int *>(sz + n);
  if (tid == 0) {
    bad    = 0;
    pivMax = 0.0;
  }
  for (int i = tid; i < n; i += kInvThreads) {
    const double *p = args.inRec + 3 * (en.inOff + i);
    sx[i] = p[0], sy[i] = p[1], sz[i] = p[2];
  }
  __syncthreads();
  // thread (tx, ty) owns the elements (tx + 16 a, ty + 8 b): consecutive threads touch consecutive rows of a column, and
  // the loops need no integer division (the flat e % n, e / n version spent most of its time on them)
  const int tx = tid & 15, ty = tid >> 4;
  // ---- buildMatrixCLU (RadialBasisFctSolver.hpp:191-192), both triangles ----
  for (int j = ty; j < n; j += kInvThreads / 16)
    for (int i = tx; i < n; i += 16)
      a[j * n + i] = pairBasis<KIND_RUNTIME, DIM3>(sx[j], sy[j], sz[j], sx[i], sy[i], sz[i], rbf, 0.0);
  __syncthreads();
  // ---- in-place Gauss-Jordan inversion with partial pivoting ----
  for (int k = 0; k < n; ++k) {
    // pivot: largest |a(i, k)|, i >= k; ties -> lowest row
    double best = -1.0;
    int    bi   = k;
    for (int i = k + tid; i < n; i += kInvThreads) {
      const double v = fabs(a[k * n + i]);
      if (v > best) {
        best = v;
        bi   = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int    oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi   = oi;
      }
    }
    if (lane == 0) {
      redVal[w] = best;
      redIdx[w] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      double b = redVal[0];
      int    r = redIdx[0];
      for (int q = 1; q < kInvThreads / 32; ++q)
        if (redVal[q] > b || (redVal[q] == b && redIdx[q] < r)) {
          b = redVal[q];
          r = redIdx[q];
        }
      pivRow  = r;
      perm[k] = r;
      pivMax  = fmax(pivMax, b);
      // rank criterion of ColPivHouseholderQR::isInvertible (threshold eps * size, relative to the largest pivot)
      if (!(b > 2.220446049250313e-16 * n * pivMax) || !(b > 0.0))
        bad = 1;
    }
    __syncthreads();
    if (bad)
      break;
    const int p = pivRow;
    // swap rows k and p, read the pivot row and the pivot column
    for (int j = tid; j < n; j += kInvThreads) {
      const double vk = a[j * n + k], vp = a[j * n + p];
      a[j * n + k]    = vp;
      a[j * n + p]    = vk;
      rowk[j]         = vp; // new row k
    }
    __syncthreads();
    for (int i = tid; i < n; i += kInvThreads)
      colk[i] = a[k * n + i]; // column k after the swap
    __syncthreads();
    const double d = 1.0 / colk[k];
    // row k <- row k / pivot with a(k, k) <- 1 / pivot; row i <- row i - a(i, k) row k with a(i, k) <- -a(i, k) / pivot
    // rank-one update of all elements with branch-free inner loops (the thread's rows of the pivot column stay in
    // registers); the pivot row and the pivot column are overwritten with their final values afterwards
    double ci[8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
      ci[q] = (tx + 16 * q < n) ? colk[tx + 16 * q] : 0.0;
    for (int j = ty; j < n; j += kInvThreads / 16) {
      const double rj  = rowk[j] * d;
      double      *col = a + j * n + tx;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (tx + 16 * q < n)
          col[16 * q] = fma(-ci[q], rj, col[16 * q]);
    }
    __syncthreads();
    for (int j = tid; j < n; j += kInvThreads) {
      a[j * n + k] = (j == k) ? d : rowk[j] * d; // row k <- row k / pivot, a(k, k) <- 1 / pivot
      if (j != k)
        a[k * n + j] = -colk[j] * d;             // column k <- -column k / pivot
    }
    __syncthreads();
  }
  if (bad) {
    if (tid == 0)
      atomicMin(fail, en.c);
    return;
  }
  // undo the row interchanges: columns of the inverse are swapped in reverse order
  for (int k = n - 1; k >= 0; --k) {
    const int p = perm[k];
    if (p != k) { // uniform
      for (int i = tid; i < n; i += kInvThreads) {
        const double vk = a[k * n + i], vp = a[p * n + i];
        a[k * n + i]    = vp;
        a[p * n + i]    = vk;
      }
      __syncthreads();
    }
  }
  double *mp = args.mat + en.matOff;
  for (int e = tid; e < n * n; e += kInvThreads)
    mp[e] = a[e]; // S[j n + i] = (C^-1)(i, j)
}

} // namespace

void pumInverse(const PumSolveArgs &a, const PumBuckets &bk, int *fail, cudaStream_t s)
{
  for (int b = 0; b < kNumBuckets; ++b) {
    const int64_t count = bk.start[b + 1] - bk.start[b];
    if (count <= 0)
      continue;
    const int    maxN = bk.maxN[b];
    const size_t sm   = (static_cast<size_t>(maxN) * maxN + 5 * static_cast<size_t>(maxN)) * sizeof(double) + static_cast<size_t>(maxN) * sizeof(int) + 16;
    if (maxN > kPumMaxInverseVertices || sm > kPumMaxSmem)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct with a basis function that is not strictly positive definite supports clusters of up to " +
                                                 std::to_string(kPumMaxInverseVertices) + " vertices (found " + std::to_string(maxN) +
                                                 "); reduce \"vertices-per-cluster\".");
    const PumEntry *list = bk.entries.p + bk.start[b];
    if (a.dim == 3) {
      if (sm > 48 * 1024)
        RBF_CUDA(cudaFuncSetAttribute(pumInverseKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm)));
      pumInverseKernel<true><<<static_cast<unsigned>(count), kInvThreads, sm, s>>>(a, list, fail);
    } else {
      if (sm > 48 * 1024)
        RBF_CUDA(cudaFuncSetAttribute(pumInverseKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm)));
      pumInverseKernel<false><<<static_cast<unsigned>(count), kInvThreads, sm, s>>>(a, list, fail);
    }
    RBF_LAUNCHED();
  }
  RBF_CUDA(cudaGetLastError());
}

} // namespace rbfmap
