// global_iterative.cu — rbf-global-iterative: matrix-free conjugate gradients on one B200.
//
// Semantics follow the reference's iterative back-ends: the same algebra as solveConsistent / solveConservative
// (RadialBasisFctSolver.hpp:413-468) with C^-1 applied by CG; stopping rules of GinkgoRadialBasisFctSolver.hpp:337-374
// (absolute 1e-30, max-iterations; the relative test uses ||b|| like the PETSc path, see cgStartKernel) and warm start from the
// coefficients of the previous call (:218-221, 437; PetRadialBasisFctMapping.hpp:452-453, 628-636 keeps one guess per
// data component, which is what is done here).
// Neither C (n x n) nor A (m x n) is ever formed (the reference's Ginkgo path materialises both densely,
// GinkgoRadialBasisFctSolver.hpp:270-271): y = Phi x is evaluated from a uniform cell grid whose cells are half a support
// radius wide, one warp per row, all vectors kept in cell order so that the gathers are contiguous.
// Functions without compact support fall back to the dense matrix-free product of dense.cu.
#include "dense.cuh"
#include "mapping.cuh"
#include "nccl_dyn.cuh"
#include "pum_device.cuh"
#include "small_dense.cuh"
#include "spatial.cuh"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>

namespace rbfmap {
namespace {

constexpr int kMaxNc = 3;

struct CgScalars {
  double rr[kMaxNc], pAp[kMaxNc], rrNew[kMaxNc], r0[kMaxNc], bb[kMaxNc];
  int    active[kMaxNc];
};

// y[c][r] = sum_j phi(|row_r - col_j|) x[c][j] over the (2H+1)^3 cell neighbourhood; rows and columns in cell order
template <int KIND, int NC, int H>
__global__ void __launch_bounds__(256) sparseApplyKernel(RbfParams rbf, double support2, BinGridView rows, long long rowBegin, long long nRows, BinGridView cols,
                                                        const double *__restrict__ x, long long ldx, double *__restrict__ y, long long ldy,
                                                        const double *__restrict__ dotWith, long long ldd, double *__restrict__ dotOut)
{
  __shared__ double dots[NC][8];
  const long long r    = rowBegin + ((blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5);
  const int       lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool      live = r < nRows;
  const long long rr   = live ? r : nRows - 1;
  const double    px = rows.xyz[3 * rr], py = rows.xyz[3 * rr + 1], pz = rows.xyz[3 * rr + 2];
  double          acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    acc[c] = 0.0;
  if (live)
  forNeighbourhoodWarp<H>(cols, px, py, pz, lane, [&](int pos, bool valid) {
    if (valid) {
      const double d2 = sqDistD<true>(px, py, pz, cols.xyz[3 * pos], cols.xyz[3 * pos + 1], cols.xyz[3 * pos + 2]);
      if (d2 <= support2) { // beyond the support every compact function evaluates to exactly 0
        const double phi = evalBasisK<KIND>(distSqrt(d2), rbf);
#pragma unroll
        for (int c = 0; c < NC; ++c)
          acc[c] = fma(phi, x[c * ldx + pos], acc[c]);
      }
    }
    return false;
  });
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (live)
        y[c * ldy + r] = acc[c];
      if (dotWith)
        dots[c][w] = live ? acc[c] * dotWith[c * ldd + r] : 0.0;
    }
  }
  if (dotWith) { // fused dot product p . (A p): one atomic per block and component
    __syncthreads();
    if (threadIdx.x < NC) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        s += dots[threadIdx.x][k];
      atomicAdd(&dotOut[threadIdx.x], s);
    }
  }
}

// ---- assembled system matrix (compactly supported functions): CSR over the cell-ordered vertices ----
// The matrix-free operator above spends its time on candidates outside the support (5x5x5 cells hold 3.7x the volume of the
// support sphere) and, because the hits are scattered over the lanes, pays the basis-function evaluation for every batch
// of 32 candidates. The system matrix C is applied once per CG iteration, hundreds of times per map: its non-zeros are
// therefore evaluated once in computeMapping and every iteration becomes a sparse matrix-vector product that streams
// 12 bytes per non-zero from HBM (what the reference's PETSc path does with its assembled SBAIJ matrix,
// PetRadialBasisFctMapping.hpp:842-1000; 1.6 GB at 500k vertices with ~260 neighbours).
template <int H>
__global__ void __launch_bounds__(256) countNeighboursKernel(double support2, BinGridView rows, long long rowBegin, long long nRows, BinGridView cols,
                                                            int *__restrict__ count)
{
  const long long r    = rowBegin + ((blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5);
  const int       lane = threadIdx.x & 31;
  if (r >= nRows)
    return;
  const double px = rows.xyz[3 * r], py = rows.xyz[3 * r + 1], pz = rows.xyz[3 * r + 2];
  int          n  = 0;
  forNeighbourhoodWarp<H>(cols, px, py, pz, lane, [&](int pos, bool valid) {
    if (valid && sqDistD<true>(px, py, pz, cols.xyz[3 * pos], cols.xyz[3 * pos + 1], cols.xyz[3 * pos + 2]) <= support2)
      ++n;
    return false;
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0)
    count[r - rowBegin] = n;
}

template <int KIND, int H>
__global__ void __launch_bounds__(256) fillMatrixKernel(RbfParams rbf, double support2, BinGridView rows, long long rowBegin, long long nRows, BinGridView cols,
                                                       const long long *__restrict__ rowOff, int *__restrict__ colIdx, double *__restrict__ val)
{
  const long long r    = rowBegin + ((blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5);
  const int       lane = threadIdx.x & 31;
  if (r >= nRows)
    return;
  const double px = rows.xyz[3 * r], py = rows.xyz[3 * r + 1], pz = rows.xyz[3 * r + 2];
  long long    at = rowOff[r - rowBegin];
  forNeighbourhoodWarp<H>(cols, px, py, pz, lane, [&](int pos, bool valid) {
    double d2  = 0.0;
    bool   hit = false;
    if (valid) {
      d2  = sqDistD<true>(px, py, pz, cols.xyz[3 * pos], cols.xyz[3 * pos + 1], cols.xyz[3 * pos + 2]);
      hit = d2 <= support2;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const long long k = at + __popc(m & ((1u << lane) - 1u));
      colIdx[k]         = pos;
      val[k]            = evalBasisK<KIND>(distSqrt(d2), rbf); // the value sparseApplyKernel computes on the fly
    }
    at += __popc(m);
    return false;
  });
}

// y[c][r] = sum_k val[k] x[c][col[k]] over the row's non-zeros; one warp per row; optional fused dot product as above
template <int NC>
__global__ void __launch_bounds__(256) spmvKernel(const long long *__restrict__ rowOff, const int *__restrict__ colIdx, const double *__restrict__ val,
                                                 long long rowBegin, long long nRows, const double *__restrict__ x, long long ldx, double *__restrict__ y,
                                                 long long ldy, const double *__restrict__ dotWith, long long ldd, double *__restrict__ dotOut)
{
  __shared__ double dots[NC][8];
  const long long r    = rowBegin + ((blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5);
  const int       lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool      live = r < nRows;
  double          acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    acc[c] = 0.0;
  if (live) {
    const long long beg = rowOff[r - rowBegin], end = rowOff[r - rowBegin + 1];
    for (long long k = beg + lane; k < end; k += 32) {
      const double v = val[k];
      const int    j = colIdx[k];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        acc[c] = fma(v, x[c * ldx + j], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (live)
        y[c * ldy + r] = acc[c];
      if (dotWith)
        dots[c][w] = live ? acc[c] * dotWith[c * ldd + r] : 0.0;
    }
  }
  if (dotWith) {
    __syncthreads();
    if (threadIdx.x < NC) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        s += dots[threadIdx.x][k];
      atomicAdd(&dotOut[threadIdx.x], s);
    }
  }
}

__global__ void gatherPlanarKernel(const double *__restrict__ in /*interleaved*/, int dims, int c0, int nc, const int *__restrict__ ids, long long n,
                                   double *__restrict__ out, long long ld)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  const long long v = ids[i];
  for (int c = 0; c < nc; ++c)
    out[c * ld + i] = in[v * dims + c0 + c];
}
__global__ void scatterInterleavedKernel(const double *__restrict__ in, long long ld, int nc, const int *__restrict__ ids, long long n, int dims, int c0,
                                         double *__restrict__ out)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  const long long v = ids[i];
  for (int c = 0; c < nc; ++c)
    out[v * dims + c0 + c] = in[c * ld + i];
}

// r = b - Ax (Ax given), p = r, rr = r.r
// (b and r may be the same array: the residual replaces the right-hand side in place, so neither is `restrict`)
template <int NC>
__global__ void cgInitKernel(const double *b, const double *__restrict__ Ax, double *r, double *__restrict__ p, long long n, long long ld,
                             CgScalars *sc)
{
  double acc[NC], accB[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    acc[c] = accB[c] = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double bv = b[c * ld + i];
      const double v  = bv - Ax[c * ld + i];
      r[c * ld + i]   = v;
      p[c * ld + i]   = v;
      acc[c] += v * v;
      accB[c] += bv * bv;
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
      accB[c] += __shfl_xor_sync(0xffffffffu, accB[c], o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&sc->rr[c], acc[c]);
      atomicAdd(&sc->bb[c], accB[c]);
    }
  }
}

// x += alpha p ; r -= alpha Ap ; rrNew = r.r   (alpha = rr / pAp per component; frozen components are skipped)
template <int NC>
__global__ void cgUpdateKernel(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p, const double *__restrict__ Ap, long long n,
                               long long ld, CgScalars *sc)
{
  double alpha[NC], acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    alpha[c] = (sc->active[c] && sc->pAp[c] != 0.0) ? sc->rr[c] / sc->pAp[c] : 0.0;
    acc[c]   = 0.0;
  }
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      x[c * ld + i] = fma(alpha[c], p[c * ld + i], x[c * ld + i]);
      const double v = fma(-alpha[c], Ap[c * ld + i], r[c * ld + i]);
      r[c * ld + i]  = v;
      acc[c] += v * v;
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if ((threadIdx.x & 31) == 0)
      atomicAdd(&sc->rrNew[c], acc[c]);
  }
}

template <int NC>
__global__ void cgDotKernel(const double *__restrict__ a, const double *__restrict__ b, long long n, long long ld, double *__restrict__ out)
{
  double acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    acc[c] = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
#pragma unroll
    for (int c = 0; c < NC; ++c)
      acc[c] += a[c * ld + i] * b[c * ld + i];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if ((threadIdx.x & 31) == 0)
      atomicAdd(&out[c], acc[c]);
  }
}

// p = r + beta p (beta = rrNew / rr), then rr <- rrNew, pAp <- 0, rrNew <- 0 by the last block
template <int NC>
__global__ void cgDirectionKernel(double *__restrict__ p, const double *__restrict__ r, long long n, long long ld, const CgScalars *sc)
{
  double beta[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    beta[c] = (sc->active[c] && sc->rr[c] != 0.0) ? sc->rrNew[c] / sc->rr[c] : 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
      p[c * ld + i] = fma(beta[c], p[c * ld + i], r[c * ld + i]);
  }
}
__global__ void cgRollKernel(CgScalars *sc, int nc, double rtol)
{
  const int c = threadIdx.x;
  if (c >= nc)
    return;
  const double res = sqrt(sc->rrNew[c]);
  if (res <= rtol * sc->r0[c] || res <= 1e-30)
    sc->active[c] = 0;
  sc->rr[c]    = sc->rrNew[c];
  sc->rrNew[c] = 0.0;
  sc->pAp[c]   = 0.0;
}
// Stopping rule: ||r|| <= solver-rtol * ||b|| (KSPConvergedDefault of the reference's PETSc path,
// PetRadialBasisFctMapping.hpp:246-263, which also starts from the previous solution) or ||r|| <= 1e-30
// (GinkgoRadialBasisFctSolver.hpp:366-369). Ginkgo's own baseline is the INITIAL residual (:352-365); with the warm
// start both back-ends use, that rule would demand a further 1/rtol reduction of an already converged residual
// whenever the data barely changes, so the right-hand-side norm is used here.
__global__ void cgStartKernel(CgScalars *sc, int nc, double rtol)
{
  const int c = threadIdx.x;
  if (c >= nc)
    return;
  sc->r0[c]     = sqrt(sc->bb[c]);
  const double res = sqrt(sc->rr[c]);
  sc->active[c] = (res <= rtol * sc->r0[c] || res <= 1e-30) ? 0 : 1;
  sc->pAp[c]    = 0.0;
  sc->rrNew[c]  = 0.0;
}

// ---- polynomial helpers (same scheme as global_direct.cu, vectors in cell order) ----
struct PolyIt {
  int    mode, mask, p;
  double cx, cy, cz, scale;
  double G[16];
};
__device__ __forceinline__ void polyRowI(const PolyIt &pg, double x, double y, double z, double (&q)[4])
{
  q[0] = 1.0;
  q[1] = q[2] = q[3] = 0.0;
  const double v[3] = {(x - pg.cx) * pg.scale, (y - pg.cy) * pg.scale, (z - pg.cz) * pg.scale};
  int          k    = 1;
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if (pg.mask & (1 << d)) {
      if (k == 1) q[1] = v[d];
      else if (k == 2) q[2] = v[d];
      else q[3] = v[d];
      ++k;
    }
}
__global__ void momentsItKernel(const double *__restrict__ xyz, long long n, double *__restrict__ m16)
{
  double m[10];
#pragma unroll
  for (int k = 0; k < 10; ++k)
    m[k] = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double q[4] = {1.0, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    int          k    = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s)
        m[k++] += q[r] * q[s];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 10; ++k)
      m[k] += __shfl_xor_sync(0xffffffffu, m[k], o);
  if ((threadIdx.x & 31) == 0) {
    int k = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s)
        atomicAdd(&m16[4 * r + s], m[k++]);
  }
}
__global__ void gramItKernel(const double *__restrict__ xyz, long long n, PolyIt pg, double *__restrict__ g16)
{
  double m[10];
#pragma unroll
  for (int k = 0; k < 10; ++k)
    m[k] = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    double q[4];
    polyRowI(pg, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], q);
    int k = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s)
        m[k++] += q[r] * q[s];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 10; ++k)
      m[k] += __shfl_xor_sync(0xffffffffu, m[k], o);
  if ((threadIdx.x & 31) == 0) {
    int k = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s)
        atomicAdd(&g16[4 * r + s], m[k++]);
  }
}
__global__ void projectItKernel(const double *__restrict__ xyz, long long n, PolyIt pg, const double *__restrict__ vals, long long ld, int nrhs,
                                const double *__restrict__ beta, double sgn, double *__restrict__ t)
{
  for (int c = 0; c < nrhs; ++c) {
    double acc[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    if (beta)
      for (int k = 0; k < 4; ++k)
        b[k] = beta[k * nrhs + c];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
      double q[4];
      polyRowI(pg, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], q);
      const double r = vals[c * ld + i] + sgn * (q[0] * b[0] + q[1] * b[1] + q[2] * b[2] + q[3] * b[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        acc[k] += q[k] * r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        atomicAdd(&t[k * nrhs + c], acc[k]);
  }
}
__global__ void solveItKernel(PolyIt pg, const double *__restrict__ a, const double *__restrict__ b, int nrhs, double *__restrict__ beta)
{
  const int c = threadIdx.x;
  if (c >= nrhs)
    return;
  double t[4];
  for (int k = 0; k < 4; ++k)
    t[k] = a[k * nrhs + c] - (b ? b[k * nrhs + c] : 0.0);
  for (int k = 0; k < 4; ++k) {
    double s = 0.0;
    for (int l = 0; l < 4; ++l)
      s += pg.G[4 * k + l] * t[l];
    beta[k * nrhs + c] += s;
  }
}
__global__ void applyItKernel(const double *__restrict__ xyz, long long n, PolyIt pg, const double *__restrict__ beta, int nrhs, double sgn,
                              double *__restrict__ vals, long long ld)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  double q[4];
  polyRowI(pg, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], q);
  for (int c = 0; c < nrhs; ++c) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      s += q[k] * beta[k * nrhs + c];
    vals[c * ld + i] += sgn * s;
  }
}
__global__ void projectAxesKernel(double *__restrict__ xyz, long long n, int dead)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  for (int d = 0; d < 3; ++d)
    if (dead & (1 << d))
      xyz[3 * i + d] = 0.0; // (u - v) * 0 of RadialBasisFctSolver.hpp:102-113
}

// VecPointwiseDivide (PETSc): w[i] = x[i] / y[i], 0 where y[i] == 0 (PetRadialBasisFctMapping.hpp:663-666)
__global__ void pointwiseDivideKernel(double *__restrict__ x, int64_t ld, int nc, const double *__restrict__ y, int64_t n)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  const double d = y[i];
  for (int c = 0; c < nc; ++c)
    x[c * ld + i] = d != 0.0 ? x[c * ld + i] / d : 0.0;
}
// VecFilter / VecChop (PetRadialBasisFctMapping.hpp:25-32): |x| < tol -> 0
__global__ void filterKernel(double *__restrict__ x, int64_t n, double tol)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n && fabs(x[i]) < tol)
    x[i] = 0.0;
}
__global__ void fillKernel(double *__restrict__ x, int64_t n, double v)
{
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n)
    x[i] = v;
}

inline int reduceGrid(int64_t n) { return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(smCount()) * 8))); }

class GlobalIterativeMapping final : public DeviceMapping {
public:
  explicit GlobalIterativeMapping(const rbfmap_config &cfg)
      : DeviceMapping(cfg)
  {
    if (!basisIsSpd(cfg.basis))
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-global-iterative (conjugate gradients) requires a strictly positive definite basis function.");
    if (cfg.polynomial == RBFMAP_POLYNOMIAL_ON)
      throw MapError(RBFMAP_ERR_INVALID, "The integrated polynomial (polynomial=\"on\") is not supported for the selected radial-basis function. "
                                         "Please select another radial-basis function or change the polynomial configuration.");
    bool anyAlive = false;
    for (int d = 0; d < cfg.dimensions; ++d)
      anyAlive |= (cfg.dead_axis[d] == 0);
    if (!anyAlive)
      throw MapError(RBFMAP_ERR_INVALID, "You cannot set all axes to dead for an RBF mapping. Please remove one of the respective mapping's "
                                         "\"x-dead\", \"y-dead\", or \"z-dead\" attributes.");
    _dead = 0;
    for (int d = 0; d < 3; ++d)
      if (d >= cfg.dimensions || cfg.dead_axis[d])
        _dead |= (1 << d);
    if (!(cfg.solver_rtol > 0.0))
      throw MapError(RBFMAP_ERR_INVALID, "solver-rtol has to be positive");
  }
  std::string name() const override { return "global-iterative RBF (cuda-executor)"; }

  void computeMapping(const double *dInputXyz, int64_t nInput, const double *dOutputXyz, int64_t nOutput) override;
  void mapConsistent(const double *dIn, int dims, double *dOut) override { map(false, dIn, dims, dOut); }
  void mapConservative(const double *dIn, int dims, double *dOut) override { map(true, dIn, dims, dOut); }
  void distributedInit(const char *uniqueId, int rank, int world) override
  {
    if (world > 1 && !uniqueId)
      throw MapError(RBFMAP_ERR_INVALID, "rbf-global-iterative needs an NCCL id to shard (all-gather / all-reduce per iteration).");
    if (_computed)
      throw MapError(RBFMAP_ERR_STATE, "Please clear the mapping before joining a group."); // the row slices depend on the group
    _group.init(uniqueId, rank, world);
    // the kept initial guess has the leading dimension of the previous world size
    _guess.release();
    _guessDims = 0;
  }
  // Mapping::initialGuess(): caller-owned warm start, one vector per data (PetRadialBasisFctMapping.hpp:176, 526-561)
  int64_t initialGuessSize() const override { return _computed ? _group.padded(_nIn) : 0; }
  void    loadInitialGuess(const double *hGuess, int dims) override
  {
    AllocStreamScope allocScope(_stream);
    const size_t     count = static_cast<size_t>(_group.padded(_nIn)) * dims;
    if (_guessDims != dims || _guess.n != count) {
      _guess.alloc(count);
      _guessDims = dims;
    }
    if (count > 0)
      RBF_CUDA(cudaMemcpyAsync(_guess.p, hGuess, count * sizeof(double), cudaMemcpyHostToDevice, _stream));
  }
  void storeInitialGuess(double *hGuess, int dims) override
  {
    const size_t count = static_cast<size_t>(_group.padded(_nIn)) * dims;
    if (count > 0 && _guessDims == dims && _guess.n == count) {
      RBF_CUDA(cudaMemcpyAsync(hGuess, _guess.p, count * sizeof(double), cudaMemcpyDeviceToHost, _stream));
      RBF_CUDA(cudaStreamSynchronize(_stream));
    }
  }
  void clear() override
  {
    _inGrid  = BinGrid();
    _outGrid = BinGrid();
    _guess.release();
    _guessDims = 0;
    _oneInterpolant.release();
    _rescale = false;
    _rowOff.release();
    _colIdx.release();
    _val.release();
    _nnz       = 0;
    _assembled = false;
    _computed  = false;
    stats      = rbfmap_stats{};
  }

private:
  void map(bool conservative, const double *dIn, int dims, double *dOut);
  // y (rows of `rows`) = Phi(rows, cols) x (cols); optional fused dot product with `dotWith`
  // rows [rb, re) only (this rank's slice); y and dotWith are indexed by the global row
  void apply(const BinGrid &rows, int64_t rb, int64_t re, const BinGrid &cols, const double *x, int64_t ldx, double *y, int64_t ldy, int nc,
             const double *dotWith, int64_t ldd, double *dotOut);
  // all ranks contribute their slice [rank * slice, ...) of every component; buffers are padded to world * slice
  void allGatherSlices(double *v, int64_t ld, int nc, int64_t slice);
  void allReduceScalars(double *d, int count);
  void cg(double *b /*in: rhs, out: destroyed*/, double *x /*in: guess, out: solution*/, int nc);
  void setupPolynomial();

  int64_t        _ldIn = 0; // leading dimension of the in-mesh vectors: world * slice
  int            _dead = 0;
  int64_t        _nIn = 0, _nOut = 0;
  bool           _sparse = true;
  double         _support = 0.0;
  BinGrid        _inGrid, _outGrid; // cell-ordered meshes (projected on the active axes)
  PolyIt         _poly{};
  DevBuf<double> _guess; // warm start, cell order, one vector per data component
  int            _guessDims = 0;
  // rescaling (PetRadialBasisFctMapping.hpp:109, 470-490): interpolant of g(x) = 1 at the output sites, cell order
  DevBuf<double> _oneInterpolant;
  bool           _rescale = false;
  void           setupRescaling();
  // assembled system matrix of this rank's row slice (CSR over cell-ordered positions), see spmvKernel
  DevBuf<int64_t> _rowOff;
  DevBuf<int>     _colIdx;
  DevBuf<double>  _val;
  int64_t         _nnz       = 0;
  bool            _assembled = false;
  void            assembleSystem();
};

void GlobalIterativeMapping::apply(const BinGrid &rows, int64_t rb, int64_t re, const BinGrid &cols, const double *x, int64_t ldx, double *y, int64_t ldy,
                                   int nc, const double *dotWith, int64_t ldd, double *dotOut)
{
  if (re <= rb)
    return;
  if (!_sparse) {
    AxisMask all{1.0, 1.0, 1.0};
    denseEvaluate(_rbf, all, rows.xyz.p + 3 * rb, re - rb, cols.xyz.p, cols.n, x, ldx, y + rb, ldy, nc, false, _stream);
    return;
  }
  const int    grid = divUp((re - rb) * 32, 256);
  const double s2   = _support * _support * (1.0 + 1e-12);
  if (_assembled && &rows == &_inGrid && &cols == &_inGrid) { // the system matrix C of this rank's rows
    const long long *off = reinterpret_cast<const long long *>(_rowOff.p);
    if (nc == 1)
      spmvKernel<1><<<grid, 256, 0, _stream>>>(off, _colIdx.p, _val.p, rb, re, x, ldx, y, ldy, dotWith, ldd, dotOut);
    else if (nc == 2)
      spmvKernel<2><<<grid, 256, 0, _stream>>>(off, _colIdx.p, _val.p, rb, re, x, ldx, y, ldy, dotWith, ldd, dotOut);
    else
      spmvKernel<3><<<grid, 256, 0, _stream>>>(off, _colIdx.p, _val.p, rb, re, x, ldx, y, ldy, dotWith, ldd, dotOut);
    RBF_LAUNCHED();
    RBF_CUDA(cudaGetLastError());
    return;
  }
  RBF_DISPATCH_HOT_BASIS(_rbf.kind, {
    if (nc == 1)
      sparseApplyKernel<KIND, 1, 2><<<grid, 256, 0, _stream>>>(_rbf, s2, rows.view(), rb, re, cols.view(), x, ldx, y, ldy, dotWith, ldd, dotOut);
    else if (nc == 2)
      sparseApplyKernel<KIND, 2, 2><<<grid, 256, 0, _stream>>>(_rbf, s2, rows.view(), rb, re, cols.view(), x, ldx, y, ldy, dotWith, ldd, dotOut);
    else
      sparseApplyKernel<KIND, 3, 2><<<grid, 256, 0, _stream>>>(_rbf, s2, rows.view(), rb, re, cols.view(), x, ldx, y, ldy, dotWith, ldd, dotOut);
  });
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

void GlobalIterativeMapping::allGatherSlices(double *v, int64_t ld, int nc, int64_t slice)
{
  if (_group.world == 1)
    return;
  const NcclApi &n = ncclApi();
  RBF_NCCL(n.GroupStart());
  for (int c = 0; c < nc; ++c)
    RBF_NCCL(n.AllGather(v + c * ld + _group.rank * slice, v + c * ld, static_cast<size_t>(slice), ncclDouble, _group.comm, _stream));
  RBF_NCCL(n.GroupEnd());
}

void GlobalIterativeMapping::allReduceScalars(double *d, int count)
{
  if (_group.world == 1)
    return;
  RBF_NCCL(ncclApi().AllReduce(d, d, static_cast<size_t>(count), ncclDouble, ncclSum, _group.comm, _stream));
}

void GlobalIterativeMapping::setupPolynomial()
{
  _poly      = PolyIt{};
  _poly.mode = 0;
  if (_cfg.polynomial != RBFMAP_POLYNOMIAL_SEPARATE)
    return;
  DevBuf<double> m16;
  m16.alloc(16);
  RBF_CUDA(cudaMemsetAsync(m16.p, 0, 16 * sizeof(double), _stream));
  momentsItKernel<<<reduceGrid(_nIn), 256, 0, _stream>>>(_inGrid.xyz.p, _nIn, m16.p);
  RBF_LAUNCHED();
  double hm[16];
  RBF_CUDA(cudaMemcpyAsync(hm, m16.p, sizeof(hm), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  double bbMin[3], bbMax[3];
  computeBounds(_inGrid.xyz.p, _nIn, bbMin, bbMax, _stream);
  int mask = 7 & ~_dead;
  int p    = 1;
  while (true) {
    p = 1 + ((mask & 1) ? 1 : 0) + ((mask & 2) ? 1 : 0) + ((mask & 4) ? 1 : 0);
    if (_nIn < p)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-global-iterative: fewer vertices than polynomial parameters are not supported on the device.");
    double g[4][4];
    for (int r = 0; r < 4; ++r)
      for (int s = 0; s < 4; ++s) {
        const bool   ar = r == 0 || (mask & (1 << (r - 1))), as = s == 0 || (mask & (1 << (s - 1)));
        const double v  = r <= s ? hm[4 * r + s] : hm[4 * s + r];
        g[r][s]         = (ar && as) ? v : (r == s ? hm[0] : 0.0);
      }
    double evMin, evMax;
    jacobiEigenvalues4(g, evMin, evMax);
    const double cond = std::sqrt(std::max(evMax, 0.0)) / std::max(std::sqrt(std::max(evMin, 0.0)), kZeroDiff);
    if (cond > 1e5) {
      int    drop = -1;
      double best = 0.0;
      for (int d = 0; d < 3; ++d)
        if (mask & (1 << d)) {
          const double e = std::abs(bbMax[d] - bbMin[d]);
          if (drop < 0 || e < best) {
            best = e;
            drop = d;
          }
        }
      if (drop < 0)
        break;
      mask &= ~(1 << drop);
    } else {
      break;
    }
  }
  _poly.mode  = 1;
  _poly.mask  = mask;
  _poly.p     = p;
  _poly.cx    = 0.5 * (bbMin[0] + bbMax[0]);
  _poly.cy    = 0.5 * (bbMin[1] + bbMax[1]);
  _poly.cz    = 0.5 * (bbMin[2] + bbMax[2]);
  double span = 0.0;
  for (int d = 0; d < 3; ++d)
    if (mask & (1 << d))
      span = std::max(span, bbMax[d] - bbMin[d]);
  _poly.scale = span > 0.0 ? 2.0 / span : 1.0;
  RBF_CUDA(cudaMemsetAsync(m16.p, 0, 16 * sizeof(double), _stream));
  gramItKernel<<<reduceGrid(_nIn), 256, 0, _stream>>>(_inGrid.xyz.p, _nIn, _poly, m16.p);
  RBF_LAUNCHED();
  RBF_CUDA(cudaMemcpyAsync(hm, m16.p, sizeof(hm), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  double g[4][4], inv[4][4];
  for (int r = 0; r < 4; ++r)
    for (int s = 0; s < 4; ++s)
      g[r][s] = r <= s ? hm[4 * r + s] : hm[4 * s + r];
  for (int r = p; r < 4; ++r)
    g[r][r] = 1.0;
  invertSpd4(g, inv);
  for (int r = 0; r < 4; ++r)
    for (int s = 0; s < 4; ++s)
      _poly.G[4 * r + s] = (r < p && s < p) ? inv[r][s] : 0.0;
}

void GlobalIterativeMapping::assembleSystem()
{
  _assembled = false;
  if (!_sparse || _nIn == 0)
    return;
  int64_t rb, re, slice;
  partitionRows(_nIn, _group.world, _group.rank, rb, re, slice);
  const int64_t m = re - rb;
  if (m <= 0) {
    _assembled = true; // no rows on this rank
    _rowOff.alloc(1);
    RBF_CUDA(cudaMemsetAsync(_rowOff.p, 0, sizeof(int64_t), _stream));
    return;
  }
  const double s2   = _support * _support * (1.0 + 1e-12);
  const int    grid = divUp(m * 32, 256);
  DevBuf<int>  count;
  count.alloc(m);
  countNeighboursKernel<2><<<grid, 256, 0, _stream>>>(s2, _inGrid.view(), rb, re, _inGrid.view(), count.p);
  RBF_LAUNCHED();
  _rowOff.alloc(m + 1);
  exclusiveScan(count.p, _rowOff.p, m, _stream);
  int64_t nnz = 0;
  RBF_CUDA(cudaMemcpyAsync(&nnz, _rowOff.p + m, sizeof(int64_t), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  // 12 bytes per non-zero; keep the matrix-free operator when the matrix would not fit comfortably
  size_t freeB = 0, totalB = 0;
  RBF_CUDA(cudaMemGetInfo(&freeB, &totalB));
  if (static_cast<double>(nnz) * 12.0 > 0.5 * static_cast<double>(freeB)) {
    _rowOff.release();
    return;
  }
  _nnz = nnz;
  _colIdx.alloc(std::max<int64_t>(nnz, 1));
  _val.alloc(std::max<int64_t>(nnz, 1));
  RBF_DISPATCH_HOT_BASIS(_rbf.kind, (fillMatrixKernel<KIND, 2><<<grid, 256, 0, _stream>>>(_rbf, s2, _inGrid.view(), rb, re, _inGrid.view(),
                                                                                        reinterpret_cast<const long long *>(_rowOff.p), _colIdx.p, _val.p)));
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
  _assembled = true;
}

// PetRadialBasisFctMapping.hpp:470-490: rescaling coefficients = C^-1 1 with the system-matrix solver, interpolant of
// g(x) = 1 at the output sites, entries below 1e-6 set to exactly 0. Only for consistent mappings with a separated
// polynomial; a solve that does not converge deactivates the rescaling (the reference warns and does the same).
void GlobalIterativeMapping::setupRescaling()
{
  _rescale = false;
  if (!_cfg.rescaling || isConservative() || _poly.mode != 1 || _nIn == 0 || _nOut == 0)
    return;
  int64_t ib, ie, islice, ob, oe, oslice;
  partitionRows(_nIn, _group.world, _group.rank, ib, ie, islice);
  partitionRows(_nOut, _group.world, _group.rank, ob, oe, oslice);
  _ldIn               = islice * _group.world;
  const int64_t ldOut = oslice * _group.world;
  DevBuf<double> rhs, coeffs;
  rhs.alloc(_ldIn);
  coeffs.alloc(_ldIn);
  fillKernel<<<divUp(_ldIn, 256), 256, 0, _stream>>>(rhs.p, _ldIn, 1.0);
  RBF_LAUNCHED();
  RBF_CUDA(cudaMemsetAsync(coeffs.p, 0, _ldIn * sizeof(double), _stream));
  try {
    cg(rhs.p, coeffs.p, 1);
  } catch (const MapError &e) {
    if (e.code != RBFMAP_ERR_NOT_CONVERGED)
      throw;
    return; // "Deactivating rescaling!"
  }
  _oneInterpolant.alloc(ldOut);
  apply(_outGrid, ob, oe, _inGrid, coeffs.p, _ldIn, _oneInterpolant.p, ldOut, 1, nullptr, 0, nullptr);
  allGatherSlices(_oneInterpolant.p, ldOut, 1, oslice);
  filterKernel<<<divUp(ldOut, 256), 256, 0, _stream>>>(_oneInterpolant.p, ldOut, 1e-6);
  RBF_LAUNCHED();
  RBF_CUDA(cudaStreamSynchronize(_stream));
  _rescale = true;
}

void GlobalIterativeMapping::computeMapping(const double *dInputXyz, int64_t nInput, const double *dOutputXyz, int64_t nOutput)
{
  g_launches = 0;
  AllocStreamScope allocScope(_stream);
  EventTimer       whole(_stream);
  whole.start();
  clear();
  _nInput  = nInput;
  _nOutput = nOutput;
  const bool cons = isConservative();
  _nIn            = cons ? nOutput : nInput;
  _nOut           = cons ? nInput : nOutput;
  stats.n_in      = _nIn;
  stats.n_out     = _nOut;
  if (_nIn == 0) {
    _computed = true;
    return;
  }
  DevBuf<double> inXyz, outXyz;
  inXyz.alloc(3 * std::max<int64_t>(_nIn, 1));
  outXyz.alloc(3 * std::max<int64_t>(_nOut, 1));
  RBF_CUDA(cudaMemcpyAsync(inXyz.p, cons ? dOutputXyz : dInputXyz, 3 * _nIn * sizeof(double), cudaMemcpyDeviceToDevice, _stream));
  if (_nOut > 0)
    RBF_CUDA(cudaMemcpyAsync(outXyz.p, cons ? dInputXyz : dOutputXyz, 3 * _nOut * sizeof(double), cudaMemcpyDeviceToDevice, _stream));
  if (_dead) {
    projectAxesKernel<<<divUp(_nIn, 256), 256, 0, _stream>>>(inXyz.p, _nIn, _dead);
    RBF_LAUNCHED();
    if (_nOut > 0) {
      projectAxesKernel<<<divUp(_nOut, 256), 256, 0, _stream>>>(outXyz.p, _nOut, _dead);
      RBF_LAUNCHED();
    }
  }
  double iMin[3], iMax[3], oMin[3] = {0, 0, 0}, oMax[3] = {0, 0, 0};
  computeBounds(inXyz.p, _nIn, iMin, iMax, _stream);
  if (_nOut > 0)
    computeBounds(outXyz.p, _nOut, oMin, oMax, _stream);
  // grid-based operator only pays off when the support is small against the domain
  _support      = basisSupportRadius(_rbf);
  double extent = 0.0;
  for (int d = 0; d < 3; ++d)
    extent = std::max(extent, iMax[d] - iMin[d]);
  _sparse = basisHasCompactSupport(_rbf.kind) && _support < 0.25 * extent;
  // cells are half a support radius wide (5x5x5 neighbourhood); the dense fallback keeps one cell per mesh
  const double cell = _sparse ? 0.5 * _support : 4.0 * std::max(extent, 1.0);
  _inGrid.build(inXyz.p, _nIn, iMin, iMax, cell, _stream);
  _outGrid.build(outXyz.p, _nOut, oMin, oMax, cell, _stream);
  setupPolynomial();
  assembleSystem();
  _computed = true; // the operator is complete; the rescaling coefficients are obtained with it
  try {
    setupRescaling();
  } catch (...) {
    _computed = false;
    throw;
  }
  stats.device_bytes       = static_cast<int64_t>(_inGrid.bytes() + _outGrid.bytes() + _oneInterpolant.bytes() + _rowOff.bytes() + _colIdx.bytes() + _val.bytes());
  stats.sum_mn             = _nnz; // non-zeros of this rank's rows of the assembled system matrix (0: matrix-free)
  stats.ms_compute_kernels = whole.stop();
  stats.kernel_launches    = g_launches;
  _computed                = true;
}

// Conjugate gradients for nc right-hand sides in lockstep. Vectors are planar with leading dimension _ldIn (padded to
// world * slice). With several ranks every rank owns the cell-ordered row slice [rb, re): it applies the operator to its
// rows (needs the full direction vector -> all-gather per iteration) and contributes partial dot products
// (-> all-reduce of 3 doubles). b: only the slice must be valid on entry; it is overwritten by the residual.
// x: full initial guess on entry (identical on all ranks), full solution on exit.
void GlobalIterativeMapping::cg(double *b, double *x, int nc)
{
  const int64_t n = _nIn, ld = _ldIn;
  int64_t       rb, re, slice;
  partitionRows(n, _group.world, _group.rank, rb, re, slice);
  const int64_t     m = re - rb;
  DevBuf<double>    p, Ap;
  DevBuf<CgScalars> sc;
  p.alloc(static_cast<size_t>(ld) * nc);
  Ap.alloc(static_cast<size_t>(ld) * nc);
  sc.alloc(1);
  RBF_CUDA(cudaMemsetAsync(sc.p, 0, sizeof(CgScalars), _stream));
  RBF_CUDA(cudaMemsetAsync(p.p, 0, static_cast<size_t>(ld) * nc * sizeof(double), _stream));
  double *scD = reinterpret_cast<double *>(sc.p); // rr[3] pAp[3] rrNew[3] r0[3] bb[3]
  // r = b - A x0 on the slice
  apply(_inGrid, rb, re, _inGrid, x, ld, Ap.p, ld, nc, nullptr, 0, nullptr);
  const int rg       = reduceGrid(std::max<int64_t>(m, 1));
  auto      launchNc = [&](auto fn) {
    if (nc == 1) fn(std::integral_constant<int, 1>{});
    else if (nc == 2) fn(std::integral_constant<int, 2>{});
    else fn(std::integral_constant<int, 3>{});
  };
  launchNc([&](auto NC) { cgInitKernel<decltype(NC)::value><<<rg, 256, 0, _stream>>>(b + rb, Ap.p + rb, b + rb, p.p + rb, m, ld, sc.p); });
  RBF_LAUNCHED();
  allReduceScalars(scD, 5 * kMaxNc);
  cgStartKernel<<<1, 32, 0, _stream>>>(sc.p, nc, _cfg.solver_rtol);
  RBF_LAUNCHED();
  allGatherSlices(p.p, ld, nc, slice);
  CgScalars h{};
  RBF_CUDA(cudaMemcpyAsync(&h, sc.p, sizeof(h), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  double *r         = b;
  int     it        = 0;
  bool    anyActive = false;
  for (int c = 0; c < nc; ++c)
    anyActive |= h.active[c] != 0;
  const int  maxIt  = _cfg.max_iterations > 0 ? _cfg.max_iterations : 1000000;
  const bool fused  = _sparse;
  double    *pApDev = scD + kMaxNc, *rrNewDev = scD + 2 * kMaxNc;
  while (anyActive && it < maxIt) {
    apply(_inGrid, rb, re, _inGrid, p.p, ld, Ap.p, ld, nc, fused ? p.p : nullptr, ld, fused ? pApDev : nullptr);
    if (!fused && m > 0) { // dense operator: separate dot product p . (A p)
      launchNc([&](auto NC) { cgDotKernel<decltype(NC)::value><<<rg, 256, 0, _stream>>>(p.p + rb, Ap.p + rb, m, ld, pApDev); });
      RBF_LAUNCHED();
    }
    allReduceScalars(pApDev, kMaxNc);
    launchNc([&](auto NC) { cgUpdateKernel<decltype(NC)::value><<<rg, 256, 0, _stream>>>(x + rb, r + rb, p.p + rb, Ap.p + rb, m, ld, sc.p); });
    allReduceScalars(rrNewDev, kMaxNc);
    launchNc([&](auto NC) { cgDirectionKernel<decltype(NC)::value><<<rg, 256, 0, _stream>>>(p.p + rb, r + rb, m, ld, sc.p); });
    allGatherSlices(p.p, ld, nc, slice);
    cgRollKernel<<<1, 32, 0, _stream>>>(sc.p, nc, _cfg.solver_rtol);
    RBF_LAUNCHED();
    RBF_LAUNCHED();
    RBF_LAUNCHED();
    ++it;
    if ((it & 7) == 0 || it < 8) { // poll the convergence flags every few iterations
      RBF_CUDA(cudaMemcpyAsync(&h, sc.p, sizeof(h), cudaMemcpyDeviceToHost, _stream));
      RBF_CUDA(cudaStreamSynchronize(_stream));
      anyActive = false;
      for (int c = 0; c < nc; ++c)
        anyActive |= h.active[c] != 0;
    }
  }
  allGatherSlices(x, ld, nc, slice);
  RBF_CUDA(cudaMemcpyAsync(&h, sc.p, sizeof(h), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  stats.iterations = std::max(stats.iterations, it);
  for (int c = 0; c < nc; ++c)
    if (h.r0[c] > 0)
      stats.residual = std::max(stats.residual, std::sqrt(h.rr[c]) / h.r0[c]);
  for (int c = 0; c < nc; ++c)
    if (h.active[c])
      throw MapError(RBFMAP_ERR_NOT_CONVERGED, "rbf-global-iterative: the conjugate-gradient solver did not reach solver-rtol within max-iterations.");
}

void GlobalIterativeMapping::map(bool conservative, const double *dIn, int dims, double *dOut)
{
  if (!_computed)
    throw MapError(RBFMAP_ERR_STATE, "The mapping has not been computed; call computeMapping() first.");
  const int        launchesBefore = g_launches;
  AllocStreamScope allocScope(_stream);
  EventTimer       t(_stream);
  t.start();
  stats.iterations   = 0;
  stats.residual     = 0.0;
  const int64_t nRes = conservative ? _nIn : _nOut;
  if (nRes > 0)
    RBF_CUDA(cudaMemsetAsync(dOut, 0, static_cast<size_t>(nRes) * dims * sizeof(double), _stream));
  if (_nIn > 0 && _nOut > 0) {
    int64_t ib, ie, islice, ob, oe, oslice;
    partitionRows(_nIn, _group.world, _group.rank, ib, ie, islice);
    partitionRows(_nOut, _group.world, _group.rank, ob, oe, oslice);
    _ldIn              = islice * _group.world;
    const int64_t ldOut = oslice * _group.world;
    if (_guessDims != dims) { // initial guess: zero on first use, kept per component afterwards
      _guess.alloc(static_cast<size_t>(_ldIn) * dims);
      RBF_CUDA(cudaMemsetAsync(_guess.p, 0, static_cast<size_t>(_ldIn) * dims * sizeof(double), _stream));
      _guessDims = dims;
    }
    for (int c0 = 0; c0 < dims; c0 += kMaxNc) {
      const int      nc = std::min(kMaxNc, dims - c0);
      double        *x  = _guess.p + static_cast<size_t>(c0) * _ldIn;
      DevBuf<double> b, o, beta, tt, eps, u;
      b.alloc(static_cast<size_t>(_ldIn) * nc);
      if (!conservative) {
        gatherPlanarKernel<<<divUp(_nIn, 256), 256, 0, _stream>>>(dIn, dims, c0, nc, _inGrid.ids.p, _nIn, b.p, _ldIn);
        RBF_LAUNCHED();
        if (_poly.mode == 1) {
          beta.alloc(4 * nc);
          tt.alloc(4 * nc);
          RBF_CUDA(cudaMemsetAsync(beta.p, 0, 4 * nc * sizeof(double), _stream));
          for (int pass = 0; pass < 2; ++pass) {
            RBF_CUDA(cudaMemsetAsync(tt.p, 0, 4 * nc * sizeof(double), _stream));
            projectItKernel<<<reduceGrid(_nIn), 256, 0, _stream>>>(_inGrid.xyz.p, _nIn, _poly, b.p, _ldIn, nc, beta.p, -1.0, tt.p);
            solveItKernel<<<1, 32, 0, _stream>>>(_poly, tt.p, nullptr, nc, beta.p);
            RBF_LAUNCHED();
            RBF_LAUNCHED();
          }
          applyItKernel<<<divUp(_nIn, 256), 256, 0, _stream>>>(_inGrid.xyz.p, _nIn, _poly, beta.p, nc, -1.0, b.p, _ldIn);
          RBF_LAUNCHED();
        }
        cg(b.p, x, nc);
        o.alloc(static_cast<size_t>(ldOut) * nc);
        apply(_outGrid, ob, oe, _inGrid, x, _ldIn, o.p, ldOut, nc, nullptr, 0, nullptr); // this rank's evaluation rows
        allGatherSlices(o.p, ldOut, nc, oslice);
        if (_rescale) { // :663-666
          pointwiseDivideKernel<<<divUp(_nOut, 256), 256, 0, _stream>>>(o.p, ldOut, nc, _oneInterpolant.p, _nOut);
          RBF_LAUNCHED();
        }
        if (_poly.mode == 1) {
          applyItKernel<<<divUp(_nOut, 256), 256, 0, _stream>>>(_outGrid.xyz.p, _nOut, _poly, beta.p, nc, 1.0, o.p, ldOut);
          RBF_LAUNCHED();
        }
        if (_cfg.output_filter > 0.0) { // :675
          filterKernel<<<divUp(ldOut * nc, 256), 256, 0, _stream>>>(o.p, ldOut * nc, _cfg.output_filter);
          RBF_LAUNCHED();
        }
        scatterInterleavedKernel<<<divUp(_nOut, 256), 256, 0, _stream>>>(o.p, ldOut, nc, _outGrid.ids.p, _nOut, dims, c0, dOut);
        RBF_LAUNCHED();
      } else {
        u.alloc(static_cast<size_t>(_nOut) * nc);
        gatherPlanarKernel<<<divUp(_nOut, 256), 256, 0, _stream>>>(dIn, dims, c0, nc, _outGrid.ids.p, _nOut, u.p, _nOut);
        RBF_LAUNCHED();
        apply(_inGrid, ib, ie, _outGrid, u.p, _nOut, b.p, _ldIn, nc, nullptr, 0, nullptr); // slice of Au = A^T u
        cg(b.p, x, nc);
        o.alloc(static_cast<size_t>(_ldIn) * nc);
        RBF_CUDA(cudaMemcpyAsync(o.p, x, static_cast<size_t>(_ldIn) * nc * sizeof(double), cudaMemcpyDeviceToDevice, _stream));
        if (_poly.mode == 1) {
          eps.alloc(4 * nc);
          tt.alloc(4 * nc);
          beta.alloc(4 * nc);
          RBF_CUDA(cudaMemsetAsync(eps.p, 0, 4 * nc * sizeof(double), _stream));
          RBF_CUDA(cudaMemsetAsync(beta.p, 0, 4 * nc * sizeof(double), _stream));
          projectItKernel<<<reduceGrid(_nOut), 256, 0, _stream>>>(_outGrid.xyz.p, _nOut, _poly, u.p, _nOut, nc, nullptr, 0.0, eps.p);
          RBF_LAUNCHED();
          for (int pass = 0; pass < 2; ++pass) {
            RBF_CUDA(cudaMemsetAsync(tt.p, 0, 4 * nc * sizeof(double), _stream));
            projectItKernel<<<reduceGrid(_nIn), 256, 0, _stream>>>(_inGrid.xyz.p, _nIn, _poly, o.p, _ldIn, nc, beta.p, 1.0, tt.p);
            solveItKernel<<<1, 32, 0, _stream>>>(_poly, eps.p, tt.p, nc, beta.p);
            RBF_LAUNCHED();
            RBF_LAUNCHED();
          }
          applyItKernel<<<divUp(_nIn, 256), 256, 0, _stream>>>(_inGrid.xyz.p, _nIn, _poly, beta.p, nc, 1.0, o.p, _ldIn);
          RBF_LAUNCHED();
        }
        if (_cfg.output_filter > 0.0) { // :799
          filterKernel<<<divUp(_ldIn * nc, 256), 256, 0, _stream>>>(o.p, _ldIn * nc, _cfg.output_filter);
          RBF_LAUNCHED();
        }
        scatterInterleavedKernel<<<divUp(_nIn, 256), 256, 0, _stream>>>(o.p, _ldIn, nc, _inGrid.ids.p, _nIn, dims, c0, dOut);
        RBF_LAUNCHED();
      }
    }
    RBF_CUDA(cudaGetLastError());
  }
  stats.ms_map_kernels = t.stop();
  stats.kernel_launches += g_launches - launchesBefore;
}

} // namespace

std::unique_ptr<DeviceMapping> makeGlobalIterativeMapping(const rbfmap_config &cfg) { return std::make_unique<GlobalIterativeMapping>(cfg); }

} // namespace rbfmap
