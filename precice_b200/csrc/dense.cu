// dense.cu — see dense.cuh. Hand-written sm_100a kernels for the global (single dense system) RBF mappings.
#include "dense.cuh"
#include "pum_device.cuh"

#include <algorithm>
#include <climits>

namespace rbfmap {
namespace {

// =============================================================================================
// assembly: 64 x 64 tile per CTA, 16 x 16 threads, 4 x 4 entries per thread
// =============================================================================================
template <int KIND, bool MASKED>
__global__ void __launch_bounds__(256) assembleKernel(RbfParams rbf, AxisMask mk, const double *__restrict__ xyz, int n, double *__restrict__ C, long long ld,
                                                      int lowerOnly)
{
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (lowerOnly && ti < tj)
    return;
  __shared__ double rx[64], ry[64], rz[64], cx[64], cy[64], cz[64];
  const int tid = threadIdx.x;
  if (tid < 128) {
    const int  loc  = tid & 63;
    const bool rows = tid < 64;
    const int  v    = (rows ? ti : tj) * 64 + loc;
    double     x = paddingCoordinate(v), y = 0.0, z = 0.0;
    if (v < n) {
      x = xyz[3 * static_cast<long long>(v)];
      y = xyz[3 * static_cast<long long>(v) + 1];
      z = xyz[3 * static_cast<long long>(v) + 2];
    }
    if (rows) {
      rx[loc] = x;
      ry[loc] = y;
      rz[loc] = z;
    } else {
      cx[loc] = x;
      cy[loc] = y;
      cz[loc] = z;
    }
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int    jl = ty * 4 + b, j = tj * 64 + jl;
    const double xj = cx[jl], yj = cy[jl], zj = cz[jl];
    double       v[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int il = tx * 4 + a, i = ti * 64 + il;
      // RadialBasisFctSolver.hpp:191-192 (u = row vertex of the upper triangle; the square is symmetric)
      const double d2 = MASKED ? sqDist3Masked(xj, yj, zj, rx[il], ry[il], rz[il], mk.mx, mk.my, mk.mz) : sqDistD<true>(xj, yj, zj, rx[il], ry[il], rz[il]);
      double       e  = evalBasisK<KIND>(distSqrt(d2), rbf);
      if (i >= n || j >= n)
        e = (i == j) ? 1.0 : 0.0;
      v[a] = e;
    }
    double *dst = C + (static_cast<long long>(ti) * 64 + tx * 4) + static_cast<long long>(j) * ld;
    *reinterpret_cast<double2 *>(dst)     = make_double2(v[0], v[1]);
    *reinterpret_cast<double2 *>(dst + 2) = make_double2(v[2], v[3]);
  }
}

// =============================================================================================
// 64 x 64 diagonal block: register-tiled Cholesky (8 x 8 threads, 8 x 8 entries each) + explicit inverse
// =============================================================================================
constexpr int kColLd = 10; // see pum_factor.cu

__global__ void __launch_bounds__(64) potrfInvKernel(double *__restrict__ C, long long ld, long long j0, double *__restrict__ inv, int *__restrict__ fail)
{
  __shared__ __align__(16) double colBuf[2][8 * kColLd];
  __shared__ double               Ls[64][65];
  __shared__ double               rdiag[64];
  const int tid = threadIdx.x, tr = tid % 8, tc = tid / 8;
  double   *blk = C + j0 + j0 * ld;

  double A[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b)
      A[a][b] = blk[(a * 8 + tr) + static_cast<long long>(b * 8 + tc) * ld];

  double rsOwn[8];
  bool   ok = true;
#pragma unroll
  for (int kb = 0; kb < 8; ++kb) {
    if (ok) {
#pragma unroll 1
      for (int kk = 0; kk < 8; ++kk) {
        double *buf = colBuf[kk & 1];
        if (tc == kk) {
#pragma unroll
          for (int a = kb; a < 8; ++a)
            buf[tr * kColLd + a] = A[a][kb];
        }
        __syncthreads();
        const double d = buf[kk * kColLd + kb];
        if (!(d > 0.0)) {
          ok = false;
          break;
        }
        const double rs = rsqrt(d);
        if (tc == kk)
          rsOwn[kb] = rs;
        double li[9], lj[9];
#pragma unroll
        for (int a2 = (kb & ~1); a2 < 8; a2 += 2) {
          const double2 u = *reinterpret_cast<const double2 *>(&buf[tr * kColLd + a2]);
          const double2 v = *reinterpret_cast<const double2 *>(&buf[tc * kColLd + a2]);
          li[a2]     = u.x * rs;
          li[a2 + 1] = u.y * rs;
          lj[a2]     = v.x * rs;
          lj[a2 + 1] = v.y * rs;
        }
        const bool colActive = tc > kk;
#pragma unroll
        for (int a = kb; a < 8; ++a) {
          if (colActive)
            A[a][kb] = fma(-li[a], lj[kb], A[a][kb]);
#pragma unroll
          for (int b = kb + 1; b <= a; ++b)
            A[a][b] = fma(-li[a], lj[b], A[a][b]);
        }
      }
    }
  }
  if (!ok) {
    if (tid == 0)
      atomicMin(fail, static_cast<int>(j0 / 64));
    return;
  }
  // L(i,j) = column value * 1/sqrt(d_j) (the diagonal owner holds d_j itself -> sqrt(d_j))
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int j = b * 8 + tc;
#pragma unroll
    for (int a = b; a < 8; ++a) {
      const int i = a * 8 + tr;
      if (i >= j) {
        const double l                                 = A[a][b] * rsOwn[b];
        blk[i + static_cast<long long>(j) * ld]        = l;
        Ls[i][j]                                       = l;
      }
    }
  }
  __syncthreads();
  rdiag[tid] = 1.0 / Ls[tid][tid];
  __syncthreads();
  // explicit inverse: thread t computes column t of X = L^-1 by forward substitution (all reads are broadcasts)
  const int t = tid;
  double    x[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    double s = (i == t) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < i; ++j)
      s = fma(-Ls[i][j], x[j], s);
    x[i] = (i >= t) ? s * rdiag[i] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 64; ++i)
    inv[i + t * 64] = x[i];
}

// =============================================================================================
// FP64 tensor-core GEMM: C(128 x BN tile) [-]= A(128 x K) * B(BN x K)^T, all column-major
// =============================================================================================
struct GemmArgs {
  const double *A;
  long long     lda;
  const double *B;
  long long     ldb;
  double       *C;
  long long     ldc;
  int           K;
  int           tilesM, tilesN;
  long long     rowOrigin, colOrigin; // global indices of C(0,0), for the lower-triangular tile test
  int           lower;                // skip tiles entirely above the diagonal
  int           subtract;             // C -= A B^T  (else C = A B^T)
  long long     validRows;            // rows of C that may be stored (row tiles may overhang)
};

__device__ __forceinline__ void cpAsync16(void *dst, const void *src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpAsyncWait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int kGemmKc = 32;

template <int BN>
__global__ void __launch_bounds__(2 * BN) dgemmNtKernel(GemmArgs g)
{
  constexpr int BM = 128, NT = 2 * BN, LDA = BM + 8, LDB = BN + 8;
  extern __shared__ __align__(16) double sm[];
  double *As = sm;                         // [2][kGemmKc][LDA]
  double *Bs = sm + 2 * kGemmKc * LDA;     // [2][kGemmKc][LDB]

  const int ti = blockIdx.x % g.tilesM, tj = blockIdx.x / g.tilesM;
  const long long row0 = static_cast<long long>(ti) * BM, col0 = static_cast<long long>(tj) * BN;
  if (g.lower && g.rowOrigin + row0 + BM - 1 < g.colOrigin + col0)
    return;
  const double *A = g.A + row0;
  const double *B = g.B + col0;
  double       *C = g.C + row0 + col0 * g.ldc;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 1, wn = warp >> 1;
  const int gq = lane >> 2, tq = lane & 3;

  auto loadChunk = [&](int buf, int k0) {
    double *as = As + buf * kGemmKc * LDA;
    double *bs = Bs + buf * kGemmKc * LDB;
#pragma unroll
    for (int p = tid; p < kGemmKc * BM / 2; p += NT) {
      const int k = p / (BM / 2), r2 = p % (BM / 2);
      cpAsync16(as + k * LDA + 2 * r2, A + 2 * r2 + static_cast<long long>(k0 + k) * g.lda);
    }
#pragma unroll
    for (int p = tid; p < kGemmKc * BN / 2; p += NT) {
      const int k = p / (BN / 2), r2 = p % (BN / 2);
      cpAsync16(bs + k * LDB + 2 * r2, B + 2 * r2 + static_cast<long long>(k0 + k) * g.ldb);
    }
    cpAsyncCommit();
  };

  double acc[8][4][2];
  if (g.subtract) {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          acc[mi][ni][e] = C[(wm * 64 + mi * 8 + gq) + static_cast<long long>(wn * 32 + ni * 8 + 2 * tq + e) * g.ldc];
  } else {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
        acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  }
  const double sign = g.subtract ? -1.0 : 1.0;

  const int nChunks = g.K / kGemmKc;
  loadChunk(0, 0);
  for (int kc = 0; kc < nChunks; ++kc) {
    if (kc + 1 < nChunks) {
      loadChunk((kc + 1) & 1, (kc + 1) * kGemmKc);
      cpAsyncWait<1>();
    } else {
      cpAsyncWait<0>();
    }
    __syncthreads();
    const double *as = As + (kc & 1) * kGemmKc * LDA + wm * 64 + gq;
    const double *bs = Bs + (kc & 1) * kGemmKc * LDB + wn * 32 + gq;
#pragma unroll
    for (int k4 = 0; k4 < kGemmKc / 4; ++k4) {
      double a[8], b[4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
        a[mi] = as[(k4 * 4 + tq) * LDA + mi * 8] * sign;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
        b[ni] = bs[(k4 * 4 + tq) * LDB + ni * 8];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    const long long r = wm * 64 + mi * 8 + gq;
    if (row0 + r < g.validRows) {
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          C[r + static_cast<long long>(wn * 32 + ni * 8 + 2 * tq + e) * g.ldc] = acc[mi][ni][e];
    }
  }
}

template <int BN>
void launchGemm(const GemmArgs &g, cudaStream_t s)
{
  if (g.tilesM <= 0 || g.tilesN <= 0)
    return;
  constexpr size_t smem = 2 * kGemmKc * (128 + 8 + BN + 8) * sizeof(double);
  static bool      configured = false;
  if (!configured) {
    RBF_CUDA(cudaFuncSetAttribute(dgemmNtKernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  dgemmNtKernel<BN><<<static_cast<unsigned>(g.tilesM) * g.tilesN, 2 * BN, smem, s>>>(g);
  RBF_LAUNCHED();
}

// =============================================================================================
// blocked triangular solves, right-hand sides planar: B[c * ldb + i]
// =============================================================================================
template <int NC>
__global__ void __launch_bounds__(256) trsvDiagFwdKernel(const double *__restrict__ L, long long ld, const double *__restrict__ invDiag, long long p0,
                                                         double *__restrict__ B, long long ldb)
{
  __shared__ double bs[NC][64], ys[NC][64];
  const int tid = threadIdx.x;
  for (int sb = 0; sb < kPanelNb / 64; ++sb) {
    const long long r0  = p0 + sb * 64;
    const double   *inv = invDiag + (r0 / 64) * 4096;
    if (tid < 64) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        bs[c][tid] = B[c * ldb + r0 + tid];
    }
    __syncthreads();
    if (tid < 64) {
      double y[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        y[c] = 0.0;
      for (int j = 0; j <= tid; ++j) { // inverse of a lower triangle is lower triangular
        const double v = inv[tid + j * 64];
#pragma unroll
        for (int c = 0; c < NC; ++c)
          y[c] = fma(v, bs[c][j], y[c]);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        ys[c][tid]               = y[c];
        B[c * ldb + r0 + tid]    = y[c];
      }
    }
    __syncthreads();
    // rows below, inside the outer block
    const long long i = r0 + 64 + tid;
    if (i < p0 + kPanelNb) {
      double acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        acc[c] = B[c * ldb + i];
      for (int j = 0; j < 64; ++j) {
        const double v = L[i + (r0 + j) * ld];
#pragma unroll
        for (int c = 0; c < NC; ++c)
          acc[c] = fma(-v, ys[c][j], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c)
        B[c * ldb + i] = acc[c];
    }
    __syncthreads();
  }
}

template <int NC>
__global__ void __launch_bounds__(256) trsvUpdateFwdKernel(const double *__restrict__ L, long long ld, long long p0, long long np, double *__restrict__ B,
                                                           long long ldb)
{
  __shared__ double ys[NC][kPanelNb];
  const int tid = threadIdx.x;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    ys[c][tid] = B[c * ldb + p0 + tid];
  __syncthreads();
  const long long i = p0 + kPanelNb + static_cast<long long>(blockIdx.x) * 256 + tid;
  if (i >= np)
    return;
  double acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    acc[c] = B[c * ldb + i];
  const double *row = L + i + p0 * ld;
#pragma unroll 4
  for (int j = 0; j < kPanelNb; ++j) {
    const double v = row[j * ld];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      acc[c] = fma(-v, ys[c][j], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < NC; ++c)
    B[c * ldb + i] = acc[c];
}

template <int NC>
__global__ void __launch_bounds__(256) trsvDiagBwdKernel(const double *__restrict__ L, long long ld, const double *__restrict__ invDiag, long long p0,
                                                         double *__restrict__ B, long long ldb)
{
  __shared__ double ts[NC][64], xs[NC][kPanelNb];
  const int tid = threadIdx.x;
  for (int sb = kPanelNb / 64 - 1; sb >= 0; --sb) {
    const long long r0  = p0 + sb * 64;
    const double   *inv = invDiag + (r0 / 64) * 4096;
    if (tid < 64) {
      // t_j = y_j - sum_{i in later sub-blocks of this outer block} L(i, j) x_i
      double t[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        t[c] = B[c * ldb + r0 + tid];
      const double *col = L + (r0 + tid) * ld;
      for (long long i = r0 + 64; i < p0 + kPanelNb; ++i) {
        const double v = col[i];
#pragma unroll
        for (int c = 0; c < NC; ++c)
          t[c] = fma(-v, xs[c][i - p0], t[c]);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c)
        ts[c][tid] = t[c];
    }
    __syncthreads();
    if (tid < 64) {
      // x = inv^T t : x_r = sum_{j >= r} inv(j, r) t_j
      double x[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        x[c] = 0.0;
      for (int j = tid; j < 64; ++j) {
        const double v = inv[j + tid * 64];
#pragma unroll
        for (int c = 0; c < NC; ++c)
          x[c] = fma(v, ts[c][j], x[c]);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        xs[c][r0 - p0 + tid]   = x[c];
        B[c * ldb + r0 + tid]  = x[c];
      }
    }
    __syncthreads();
  }
}

// y_j -= sum_{i in [p0, p0+256)} L(i, j) x_i  for all columns j < p0: one warp per column
template <int NC>
__global__ void __launch_bounds__(256) trsvUpdateBwdKernel(const double *__restrict__ L, long long ld, long long p0, double *__restrict__ B, long long ldb)
{
  __shared__ double xs[NC][kPanelNb];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    xs[c][tid] = B[c * ldb + p0 + tid];
  __syncthreads();
  const long long j = static_cast<long long>(blockIdx.x) * 8 + w;
  if (j >= p0)
    return;
  const double *col = L + p0 + j * ld;
  double        acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    acc[c] = 0.0;
#pragma unroll
  for (int q = 0; q < kPanelNb / 32; ++q) {
    const double v = col[q * 32 + lane];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      acc[c] = fma(v, xs[c][q * 32 + lane], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
      B[c * ldb + j] -= acc[c];
  }
}

template <int NC>
void choleskySolveNc(const double *L, int64_t np, const double *invDiag, double *B, int64_t ldb, cudaStream_t s)
{
  const int64_t nPanels = np / kPanelNb;
  for (int64_t P = 0; P < nPanels; ++P) {
    const long long p0 = P * kPanelNb;
    trsvDiagFwdKernel<NC><<<1, 256, 0, s>>>(L, np, invDiag, p0, B, ldb);
    RBF_LAUNCHED();
    const int64_t rest = np - p0 - kPanelNb;
    if (rest > 0) {
      trsvUpdateFwdKernel<NC><<<divUp(rest, 256), 256, 0, s>>>(L, np, p0, np, B, ldb);
      RBF_LAUNCHED();
    }
  }
  for (int64_t P = nPanels - 1; P >= 0; --P) {
    const long long p0 = P * kPanelNb;
    trsvDiagBwdKernel<NC><<<1, 256, 0, s>>>(L, np, invDiag, p0, B, ldb);
    RBF_LAUNCHED();
    if (p0 > 0) {
      trsvUpdateBwdKernel<NC><<<divUp(p0, 8), 256, 0, s>>>(L, np, p0, B, ldb);
      RBF_LAUNCHED();
    }
  }
  RBF_CUDA(cudaGetLastError());
}

// =============================================================================================
// matrix-free product with the evaluation matrix: one thread per row, column tiles staged in shared memory
// =============================================================================================
template <int KIND, int NC, bool MASKED>
__global__ void __launch_bounds__(128) evalDenseKernel(RbfParams rbf, AxisMask mk, const double *__restrict__ rowXyz, long long nRows,
                                                       const double *__restrict__ colXyz, long long nCols, const double *__restrict__ colVals, long long ldv,
                                                       double *__restrict__ rowOut, long long ldo, int accumulate)
{
  constexpr int LDV = NC == 1 ? 4 : 6;
  __shared__ __align__(16) double tile[128 * LDV];
  const int       tid = threadIdx.x;
  const long long r   = static_cast<long long>(blockIdx.x) * 128 + tid;
  double          x = 0, y = 0, z = 0;
  if (r < nRows) {
    x = rowXyz[3 * r];
    y = rowXyz[3 * r + 1];
    z = rowXyz[3 * r + 2];
  }
  double acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c)
    acc[c] = 0.0;
  for (long long c0 = 0; c0 < nCols; c0 += 128) {
    __syncthreads();
    const long long j = c0 + tid;
    if (j < nCols) {
      tile[tid * LDV]     = colXyz[3 * j];
      tile[tid * LDV + 1] = colXyz[3 * j + 1];
      tile[tid * LDV + 2] = colXyz[3 * j + 2];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        tile[tid * LDV + 3 + c] = colVals[c * ldv + j];
    }
    __syncthreads();
    const int     cnt = static_cast<int>(min(128LL, nCols - c0));
    const double *vp  = tile;
#pragma unroll 2
    for (int jj = 0; jj < cnt; ++jj, vp += LDV) {
      const double2 xy = *reinterpret_cast<const double2 *>(vp);
      const double2 zb = *reinterpret_cast<const double2 *>(vp + 2);
      // RadialBasisFctSolver.hpp:229-236: u = row (evaluation) vertex, v = column (basis) vertex
      const double d2  = MASKED ? sqDist3Masked(x, y, z, xy.x, xy.y, zb.x, mk.mx, mk.my, mk.mz) : sqDistD<true>(x, y, z, xy.x, xy.y, zb.x);
      const double phi = evalBasisK<KIND>(distSqrt(d2), rbf);
      acc[0]           = fma(phi, zb.y, acc[0]);
      if (NC > 1) {
        const double2 b12 = *reinterpret_cast<const double2 *>(vp + 4);
        acc[1]            = fma(phi, b12.x, acc[1]);
        if (NC > 2)
          acc[2] = fma(phi, b12.y, acc[2]);
      }
    }
  }
  if (r < nRows) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double *o = rowOut + c * ldo + r;
      *o        = accumulate ? *o + acc[c] : acc[c];
    }
  }
}

template <int KIND, int NC>
void evalLaunch(const RbfParams &rbf, const AxisMask &mask, const double *rowXyz, int64_t nRows, const double *colXyz, int64_t nCols, const double *colVals,
                int64_t ldv, double *rowOut, int64_t ldo, bool accumulate, cudaStream_t s)
{
  const int grid = divUp(nRows, 128);
  if (mask.allActive())
    evalDenseKernel<KIND, NC, false><<<grid, 128, 0, s>>>(rbf, mask, rowXyz, nRows, colXyz, nCols, colVals, ldv, rowOut, ldo, accumulate ? 1 : 0);
  else
    evalDenseKernel<KIND, NC, true><<<grid, 128, 0, s>>>(rbf, mask, rowXyz, nRows, colXyz, nCols, colVals, ldv, rowOut, ldo, accumulate ? 1 : 0);
  RBF_LAUNCHED();
}

__global__ void interleavedToPlanarKernel(const double *__restrict__ in, long long n, int dims, double *__restrict__ out, long long ld, long long nFill)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= nFill)
    return;
  for (int c = 0; c < dims; ++c)
    out[c * ld + i] = i < n ? in[i * dims + c] : 0.0;
}
__global__ void planarToInterleavedKernel(const double *__restrict__ in, long long ld, long long n, int dims, double *__restrict__ out)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  for (int c = 0; c < dims; ++c)
    out[i * dims + c] = in[c * ld + i];
}

} // namespace

// =============================================================================================
void denseAssemble(const RbfParams &rbf, const AxisMask &mask, const double *dXyz, int64_t n, double *dC, int64_t np, bool lowerOnly, cudaStream_t s)
{
  const dim3 grid(static_cast<unsigned>(np / 64), static_cast<unsigned>(np / 64));
  RBF_DISPATCH_HOT_BASIS(rbf.kind, {
    if (mask.allActive())
      assembleKernel<KIND, false><<<grid, 256, 0, s>>>(rbf, mask, dXyz, static_cast<int>(n), dC, np, lowerOnly ? 1 : 0);
    else
      assembleKernel<KIND, true><<<grid, 256, 0, s>>>(rbf, mask, dXyz, static_cast<int>(n), dC, np, lowerOnly ? 1 : 0);
  });
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

bool denseCholesky(double *dC, int64_t np, double *dInvDiag, cudaStream_t s)
{
  DevBuf<int> fail;
  fail.alloc(1);
  const int noFail = INT_MAX;
  RBF_CUDA(cudaMemcpyAsync(fail.p, &noFail, sizeof(int), cudaMemcpyHostToDevice, s));
  for (int64_t p0 = 0; p0 < np; p0 += kPanelNb) {
    for (int sb = 0; sb < kPanelNb / kDiagNb; ++sb) {
      const int64_t j0 = p0 + sb * kDiagNb;
      double       *inv = dInvDiag + (j0 / kDiagNb) * kDiagNb * kDiagNb;
      potrfInvKernel<<<1, 64, 0, s>>>(dC, np, j0, inv, fail.p);
      RBF_LAUNCHED();
      const int64_t r0 = j0 + kDiagNb; // first row below the diagonal block
      if (r0 >= np)
        continue;
      // panel: L(r0:, j0:j0+64) = A(r0:, j0:j0+64) * inv(L11)^T     (in place)
      GemmArgs g{};
      g.A = dC + r0 + j0 * np;
      g.lda = np;
      g.B = inv;
      g.ldb = kDiagNb;
      g.C = dC + r0 + j0 * np;
      g.ldc = np;
      g.K = kDiagNb;
      g.tilesM = divUp(np - r0, 128);
      g.tilesN = 1;
      g.rowOrigin = r0;
      g.colOrigin = j0;
      g.lower = 0;
      g.subtract = 0;
      g.validRows = np - r0;
      launchGemm<64>(g, s);
      // update of the remaining columns of the outer panel: A(i, jc) -= L(i, j0 block) L(jc, j0 block)^T
      const int64_t cEnd = p0 + kPanelNb;
      if (r0 < cEnd) {
        GemmArgs u{};
        u.A = dC + r0 + j0 * np;
        u.lda = np;
        u.B = dC + r0 + j0 * np;
        u.ldb = np;
        u.C = dC + r0 + r0 * np;
        u.ldc = np;
        u.K = kDiagNb;
        u.tilesM = divUp(np - r0, 128);
        u.tilesN = static_cast<int>((cEnd - r0) / 64);
        u.rowOrigin = r0;
        u.colOrigin = r0;
        u.lower = 1;
        u.subtract = 1;
        u.validRows = np - r0;
        launchGemm<64>(u, s);
      }
    }
    // trailing update with the whole outer panel: K = 256
    const int64_t t0 = p0 + kPanelNb;
    if (t0 < np) {
      GemmArgs t{};
      t.A = dC + t0 + p0 * np;
      t.lda = np;
      t.B = dC + t0 + p0 * np;
      t.ldb = np;
      t.C = dC + t0 + t0 * np;
      t.ldc = np;
      t.K = kPanelNb;
      t.tilesM = static_cast<int>((np - t0) / 128);
      t.tilesN = static_cast<int>((np - t0) / 128);
      t.rowOrigin = t0;
      t.colOrigin = t0;
      t.lower = 1;
      t.subtract = 1;
      t.validRows = np - t0;
      launchGemm<128>(t, s);
    }
  }
  RBF_CUDA(cudaGetLastError());
  int failed = INT_MAX;
  RBF_CUDA(cudaMemcpyAsync(&failed, fail.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  RBF_CUDA(cudaStreamSynchronize(s));
  return failed == INT_MAX;
}

// C(rows x cols) -= A(rows x K) * B(cols x K)^T ; rows, cols multiples of 64, K multiple of 32
void denseGemmNtSubtract(const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int64_t rows, int64_t cols, int K, cudaStream_t s)
{
  GemmArgs g{};
  g.A = A;
  g.lda = lda;
  g.B = B;
  g.ldb = ldb;
  g.C = C;
  g.ldc = ldc;
  g.K = K;
  g.tilesM = divUp(rows, 128);
  g.tilesN = static_cast<int>(cols / 64);
  g.rowOrigin = 0;
  g.colOrigin = 0;
  g.lower = 0;
  g.subtract = 1;
  g.validRows = rows;
  launchGemm<64>(g, s);
}

void denseCholeskySolve(const double *dL, int64_t np, const double *dInvDiag, double *dB, int64_t ldb, int nrhs, cudaStream_t s)
{
  for (int c0 = 0; c0 < nrhs;) {
    const int nc = std::min(3, nrhs - c0);
    double   *b  = dB + c0 * ldb;
    if (nc == 1)
      choleskySolveNc<1>(dL, np, dInvDiag, b, ldb, s);
    else if (nc == 2)
      choleskySolveNc<2>(dL, np, dInvDiag, b, ldb, s);
    else
      choleskySolveNc<3>(dL, np, dInvDiag, b, ldb, s);
    c0 += nc;
  }
}

void denseEvaluate(const RbfParams &rbf, const AxisMask &mask, const double *dRowXyz, int64_t nRows, const double *dColXyz, int64_t nCols,
                   const double *dColVals, int64_t ldv, double *dRowOut, int64_t ldo, int nrhs, bool accumulate, cudaStream_t s)
{
  if (nRows <= 0)
    return;
  for (int c0 = 0; c0 < nrhs;) {
    const int     nc = std::min(3, nrhs - c0);
    const double *v  = dColVals + c0 * ldv;
    double       *o  = dRowOut + c0 * ldo;
    RBF_DISPATCH_HOT_BASIS(rbf.kind, {
      if (nc == 1)
        evalLaunch<KIND, 1>(rbf, mask, dRowXyz, nRows, dColXyz, nCols, v, ldv, o, ldo, accumulate, s);
      else if (nc == 2)
        evalLaunch<KIND, 2>(rbf, mask, dRowXyz, nRows, dColXyz, nCols, v, ldv, o, ldo, accumulate, s);
      else
        evalLaunch<KIND, 3>(rbf, mask, dRowXyz, nRows, dColXyz, nCols, v, ldv, o, ldo, accumulate, s);
    });
    c0 += nc;
  }
  RBF_CUDA(cudaGetLastError());
}

void interleavedToPlanar(const double *dIn, int64_t n, int dims, double *dOut, int64_t ld, int64_t nFill, cudaStream_t s)
{
  if (nFill <= 0)
    return;
  interleavedToPlanarKernel<<<divUp(nFill, 256), 256, 0, s>>>(dIn, n, dims, dOut, ld, nFill);
  RBF_LAUNCHED();
}
void planarToInterleaved(const double *dIn, int64_t ld, int64_t n, int dims, double *dOut, cudaStream_t s)
{
  if (n <= 0)
    return;
  planarToInterleavedKernel<<<divUp(n, 256), 256, 0, s>>>(dIn, ld, n, dims, dOut);
  RBF_LAUNCHED();
}

} // namespace rbfmap
