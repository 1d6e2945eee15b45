// pum_solve.cu — per-cluster dense algebra of the partition-of-unity mapping, hand-written for sm_100a.
//
// One CTA per cluster. Replaces, per cluster,
//   RadialBasisFctSolver.hpp:164-207 buildMatrixCLU  + :352-353 Eigen::LLT        -> pumFactorKernel
//   RadialBasisFctSolver.hpp:377-410 separated polynomial (Q, cond check, QR)       -> polySetupKernel
//   RadialBasisFctSolver.hpp:209-242 buildMatrixA + :441-468 solveConsistent        -> pumMapConsistentKernel
//   RadialBasisFctSolver.hpp:413-438 solveConservative                              -> pumMapConservativeKernel
//   SphericalVertexCluster.hpp:201-240, 296-335 gather / weight / accumulate
// (Kokkos prior art: device/KokkosPUMKernels_Impl.hpp:305-380 assembly, :458-485 LU, :491-814 solve.)
//
// Factorisation: C = M D^-1 M^T, the square-root-free form of the Cholesky factorisation
// (M = L diag(L_kk), D = diag(L_kk^2)). The pivots d_k are exactly the squares of Eigen's LLT
// pivots, so "d_k <= 0" is the same failure criterion as LLT's info() != Success.
// Packed storage per cluster (column-major lower triangle, n(n+1)/2 doubles): column k holds
// [1/d_k, M(k+1,k), ..., M(n-1,k)].
#include "pum.cuh"

#include <algorithm>
#include <climits>

namespace rbfmap {
namespace {

// =============================================================================================
// small helpers
// =============================================================================================
__device__ __forceinline__ long long colStart(int k, int n) { return static_cast<long long>(k) * n - static_cast<long long>(k) * (k - 1) / 2; }

__device__ __forceinline__ void maskToAxes(int mask, double &mx, double &my, double &mz)
{
  mx = (mask & 1) ? 1.0 : 0.0;
  my = (mask & 2) ? 1.0 : 0.0;
  mz = (mask & 4) ? 1.0 : 0.0;
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

// eigenvalues of a symmetric 4x4 matrix (cyclic Jacobi, everything in registers)
__device__ void jacobiEigenvalues4(double a[4][4], double &evMin, double &evMax)
{
#pragma unroll 1
  for (int sweep = 0; sweep < 16; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q)
        off += a[p][q] * a[p][q];
    if (off == 0.0)
      break;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
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
#pragma unroll
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

// Column-pivoted Householder QR of a tiny rows x cols matrix (rows <= 3, cols <= 4), same pivoting and
// rank rules as Eigen::ColPivHouseholderQR, condensed into the operator S (cols x rows) with
// solve(b) = S b  and  transpose().solve(c) = S^T c. Only used for clusters with fewer vertices than
// polynomial parameters (SphericalVertexCluster.hpp:170-172).
__device__ void tinyColPivQrOperator(const double Qin[3][4], int rows, int cols, double S[4][3])
{
  double A[3][4], tau[3], nUpd[4], nDir[4];
  int    perm[4];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      A[i][j] = (i < rows && j < cols) ? Qin[i][j] : 0.0;
  const int size = min(rows, cols);
  double    maxNorm = 0.0;
  for (int j = 0; j < cols; ++j) {
    double s = 0.0;
    for (int i = 0; i < rows; ++i)
      s += A[i][j] * A[i][j];
    nUpd[j] = nDir[j] = sqrt(s);
    perm[j]           = j;
    maxNorm           = fmax(maxNorm, nUpd[j]);
  }
  const double eps    = 2.220446049250313e-16;
  const double helper = (maxNorm * eps) * (maxNorm * eps) / static_cast<double>(rows);
  const double dthr   = sqrt(eps);
  int          np     = size;
  for (int k = 0; k < size; ++k) {
    int    big  = k;
    double best = nUpd[k];
    for (int j = k + 1; j < cols; ++j)
      if (nUpd[j] > best) {
        best = nUpd[j];
        big  = j;
      }
    if (np == size && best * best < helper * static_cast<double>(rows - k))
      np = k;
    if (big != k) {
      for (int i = 0; i < rows; ++i) {
        const double t = A[i][k];
        A[i][k]        = A[i][big];
        A[i][big]      = t;
      }
      double t  = nUpd[k];
      nUpd[k]   = nUpd[big];
      nUpd[big] = t;
      t         = nDir[k];
      nDir[k]   = nDir[big];
      nDir[big] = t;
      const int pi = perm[k];
      perm[k]      = perm[big];
      perm[big]    = pi;
    }
    double tailSq = 0.0;
    for (int i = k + 1; i < rows; ++i)
      tailSq += A[i][k] * A[i][k];
    const double c0 = A[k][k];
    double       beta, t;
    if (tailSq <= 2.2250738585072014e-308) {
      t    = 0.0;
      beta = c0;
      for (int i = k + 1; i < rows; ++i)
        A[i][k] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tailSq);
      if (c0 >= 0.0)
        beta = -beta;
      for (int i = k + 1; i < rows; ++i)
        A[i][k] /= (c0 - beta);
      t = (beta - c0) / beta;
    }
    tau[k]  = t;
    A[k][k] = beta;
    if (t != 0.0)
      for (int j = k + 1; j < cols; ++j) {
        double s = A[k][j];
        for (int i = k + 1; i < rows; ++i)
          s += A[i][k] * A[i][j];
        s *= t;
        A[k][j] -= s;
        for (int i = k + 1; i < rows; ++i)
          A[i][j] -= s * A[i][k];
      }
    for (int j = k + 1; j < cols; ++j)
      if (nUpd[j] != 0.0) {
        double tmp = fabs(A[k][j]) / nUpd[j];
        tmp        = (1.0 + tmp) * (1.0 - tmp);
        tmp        = tmp < 0.0 ? 0.0 : tmp;
        const double r2 = nUpd[j] / nDir[j];
        if (tmp * r2 * r2 <= dthr) {
          double s = 0.0;
          for (int i = k + 1; i < rows; ++i)
            s += A[i][j] * A[i][j];
          nDir[j] = sqrt(s);
          nUpd[j] = nDir[j];
        } else {
          nUpd[j] *= sqrt(tmp);
        }
      }
  }
  for (int q = 0; q < 4; ++q)
    for (int i = 0; i < 3; ++i)
      S[q][i] = 0.0;
  for (int e = 0; e < rows; ++e) {
    double c[3] = {0.0, 0.0, 0.0};
    c[e]        = 1.0;
    for (int k = 0; k < np; ++k) { // c := Q^T c
      if (tau[k] == 0.0)
        continue;
      double s = c[k];
      for (int i = k + 1; i < rows; ++i)
        s += A[i][k] * c[i];
      s *= tau[k];
      c[k] -= s;
      for (int i = k + 1; i < rows; ++i)
        c[i] -= s * A[i][k];
    }
    for (int i = np - 1; i >= 0; --i) {
      double s = c[i];
      for (int k = i + 1; k < np; ++k)
        s -= A[i][k] * c[k];
      c[i] = s / A[i][i];
    }
    for (int i = 0; i < np; ++i)
      S[perm[i]][e] = c[i];
  }
}

// =============================================================================================
// polynomial set-up: one thread per cluster
// =============================================================================================
__global__ void polySetupKernel(PumSolveArgs a, int64_t nClusters)
{
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c >= nClusters)
    return;
  PolyRec     &rec = a.poly[c];
  const int    n   = a.nIn[c];
  const int   *ids = a.inIds + a.inOff[c];
  // SphericalVertexCluster.hpp:178: deadAxis = vector<bool>(dim, false) -> z is inactive for 2-D meshes
  int mask = a.dim == 3 ? 7 : 3;

  // axis extents (reduceActiveAxis, RadialBasisFctSolver.hpp:115-137) and the first three vertices
  double lo[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
  double hi[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
  double first[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d)
      first[i][d] = 0.0;
  // raw moments sum q q^T with q = [1, x, y, z]
  double m[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      m[i][j] = 0.0;
  for (int i = 0; i < n; ++i) {
    const double *x    = a.inXyz + 3 * static_cast<long long>(ids[i]);
    const double  q[4] = {1.0, x[0], x[1], x[2]};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = fmin(lo[d], x[d]);
      hi[d] = fmax(hi[d], x[d]);
      if (i < 3) {
        if (i == 0) first[0][d] = x[d];
        if (i == 1) first[1][d] = x[d];
        if (i == 2) first[2][d] = x[d];
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = r; s < 4; ++s)
        m[r][s] += q[r] * q[s];
  }

  // condition-number loop (RadialBasisFctSolver.hpp:383-402): sigma_max / max(sigma_min, 1e-14) > 1e5 -> drop the
  // active axis with the smallest extent. Singular values come from the Gram matrix of Q (or of Q^T when n < p).
  int p = 1;
  while (true) {
    p = 1 + ((mask & 1) ? 1 : 0) + ((mask & 2) ? 1 : 0) + ((mask & 4) ? 1 : 0);
    double g[4][4];
    if (n >= p) {
      // select active rows/cols of the moment matrix; inactive slots become decoupled copies of g00
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const bool ar = r == 0 || (mask & (1 << (r - 1))), as = s == 0 || (mask & (1 << (s - 1)));
          const double v = r <= s ? m[r][s] : m[s][r];
          g[r][s]        = (ar && as) ? v : (r == s ? m[0][0] : 0.0);
        }
    } else {
      // n x n Gram matrix of the rows (n <= 3)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          double v = 0.0;
          if (r < 3 && s < 3) {
            v = 1.0;
#pragma unroll
            for (int d = 0; d < 3; ++d)
              if (mask & (1 << d))
                v += first[r][d] * first[s][d];
          }
          g[r][s] = v;
        }
      const double g00 = g[0][0];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (r >= n || s >= n)
            g[r][s] = (r == s) ? g00 : 0.0;
    }
    double evMin, evMax;
    jacobiEigenvalues4(g, evMin, evMax);
    const double sMax = sqrt(fmax(evMax, 0.0)), sMin = sqrt(fmax(evMin, 0.0));
    const double cond = sMax / fmax(sMin, kZeroDiff);
    if (cond > 1e5) {
      // smallest extent among the active axes, ties -> lowest axis (stable insertion sort of three pairs)
      int    drop = 0;
      double best = 1.7976931348623157e308;
      bool   any  = false;
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (mask & (1 << d)) {
          const double e = fabs(hi[d] - lo[d]);
          if (!any || e < best) {
            best = e;
            drop = d;
            any  = true;
          }
        }
      if (!any)
        break; // only the constant is left; cond == 1 in exact arithmetic
      mask &= ~(1 << drop);
    } else {
      break;
    }
  }
  rec.mask = mask;
  rec.p    = p;
  rec.pad  = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    rec.G[i] = 0.0;

  if (n >= p) {
    rec.mode = 1;
    // Gram matrix of the centred/scaled basis [1, (x_a - c_a)/r] over the active axes, then its inverse
    const double cx = a.centers[3 * c], cy = a.centers[3 * c + 1], cz = a.centers[3 * c + 2];
    const double rinv = 1.0 / a.radius;
    double       g[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s)
        g[r][s] = 0.0;
    for (int i = 0; i < n; ++i) {
      const double *x = a.inXyz + 3 * static_cast<long long>(ids[i]);
      double        q[4];
      polyRow(1, mask, x[0], x[1], x[2], cx, cy, cz, rinv, q);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s)
          g[r][s] += q[r] * q[s];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r >= p)
        g[r][r] = 1.0;
    // Gauss-Jordan inverse of the SPD 4x4 (no pivoting needed)
    double inv[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s)
        inv[r][s] = r == s ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double piv = 1.0 / g[k][k];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        g[k][s] *= piv;
        inv[k][s] *= piv;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r != k) {
          const double f = g[r][k];
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            g[r][s] -= f * g[k][s];
            inv[r][s] -= f * inv[k][s];
          }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s)
        rec.G[4 * r + s] = (r < p && s < p) ? inv[r][s] : 0.0;
  } else {
    rec.mode = 2;
    double Q[3][4], S[4][3];
    for (int i = 0; i < 3; ++i) {
      Q[i][0] = 1.0;
      int k   = 1;
      for (int d = 0; d < 3; ++d)
        if (mask & (1 << d))
          Q[i][k++] = first[i][d];
      for (; k < 4; ++k)
        Q[i][k] = 0.0;
    }
    tinyColPivQrOperator(Q, n, p, S);
    for (int q = 0; q < 4; ++q)
      for (int i = 0; i < 3; ++i)
        rec.G[4 * q + i] = S[q][i];
  }
}

__global__ void polyOffKernel(PolyRec *poly, int64_t nClusters)
{
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c < nClusters) {
    poly[c].mode = 0;
    poly[c].mask = 0;
    poly[c].p    = 0;
    poly[c].pad  = 0;
  }
}

// =============================================================================================
// assembly + factorisation, register-tiled: TG x TG threads, 8 x 8 entries per thread, n <= 8*TG
// =============================================================================================
// Element (i, j) = (a*TG + tr, b*TG + tc) lives in A[a][b] of thread (tr, tc) (block-cyclic), so the active
// part of the trailing matrix stays evenly spread over the threads while the factorisation proceeds.
template <int KIND, int TG>
__global__ void __launch_bounds__(TG *TG) pumFactorKernel(PumSolveArgs args, const int *__restrict__ list, int *__restrict__ fail)
{
  constexpr int NMAX = 8 * TG;
  __shared__ double sx[NMAX], sy[NMAX], sz[NMAX];
  __shared__ __align__(16) double colBuf[2][NMAX];

  const int c   = list[blockIdx.x];
  const int n   = args.nIn[c];
  const int tid = threadIdx.x;
  const int tr = tid % TG, tc = tid / TG;
  const int      *ids = args.inIds + args.inOff[c];
  double         *mp  = args.mat + args.matOff[c];
  const RbfParams rbf = args.rbf;

  for (int i = tid; i < NMAX; i += TG * TG) {
    double x = 0.0, y = 0.0, z = 0.0;
    if (i < n) {
      const double *p = args.inXyz + 3 * static_cast<long long>(ids[i]);
      x = p[0];
      y = p[1];
      z = p[2];
    }
    sx[i] = x;
    sy[i] = y;
    sz[i] = z;
  }
  __syncthreads();

  // ---- assembly of the lower block triangle (buildMatrixCLU) ----
  const double mz = args.dim == 3 ? 1.0 : 0.0;
  double       A[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int  i    = a * TG + tr;
    const bool rowOk = a * TG < n; // warp-uniform
    double     xi = 0, yi = 0, zi = 0;
    if (rowOk) {
      xi = sx[i];
      yi = sy[i];
      zi = sz[i];
    }
#pragma unroll
    for (int b = 0; b <= a; ++b) {
      double v = 0.0;
      if (rowOk) {
        const int j = b * TG + tc;
        if (i < n && j < n) {
          // RadialBasisFctSolver.hpp:191-192: phi(sqrt(squaredDifference(u, v, activeAxis)))
          const double d2 = sqDist3Masked(sx[j], sy[j], sz[j], xi, yi, zi, 1.0, 1.0, mz);
          v               = evalBasisT<KIND>(__dsqrt_rn(d2), rbf);
        } else if (i == j) {
          v = 1.0;
        }
      }
      A[a][b] = v;
    }
  }

  // ---- right-looking factorisation, one column per step ----
  bool ok = true;
#pragma unroll
  for (int kb = 0; kb < 8; ++kb) {
    if (kb * TG < n && ok) {
      const int kend = min(TG, n - kb * TG);
#pragma unroll 1
      for (int kk = 0; kk < kend; ++kk) {
        const int k   = kb * TG + kk;
        double   *buf = colBuf[k & 1];
        if (tc == kk) {
#pragma unroll
          for (int a = kb; a < 8; ++a)
            if (a * TG < n)
              buf[tr * 8 + a] = A[a][kb];
        }
        __syncthreads();
        const double d = buf[kk * 8 + kb];
        if (!(d > 0.0)) { // same pivot test as Eigen::LLT (x <= 0 -> NumericalIssue)
          ok = false;
          break;
        }
        const double rd = 1.0 / d;
        // column k of the packed factor: [1/d, M(k+1..n-1, k)]
        {
          const long long cs = colStart(k, n);
          for (int t = tid; t < n - k; t += TG * TG) {
            const int i = k + t;
            mp[cs + t]  = t == 0 ? rd : buf[(i % TG) * 8 + i / TG];
          }
        }
        double li[8], lj[8];
#pragma unroll
        for (int a = kb; a < 8; ++a) {
          li[a] = buf[tr * 8 + a];
          lj[a] = buf[tc * 8 + a] * rd;
        }
#pragma unroll
        for (int a = kb; a < 8; ++a)
          if (a * TG < n) {
#pragma unroll
            for (int b = kb; b <= a; ++b) {
              if (b > kb || tc > kk)
                A[a][b] = fma(-li[a], lj[b], A[a][b]);
            }
          }
      }
    }
  }
  if (!ok && tid == 0)
    atomicMin(fail, c);
}

// ---- generic variant for n > 128: factor in place in the packed global storage ----
template <int KIND>
__global__ void __launch_bounds__(256) pumFactorGenericKernel(PumSolveArgs args, const int *__restrict__ list, int *__restrict__ fail)
{
  extern __shared__ double dyn[]; // sx, sy, sz, col : 4 n doubles
  const int       c   = list[blockIdx.x];
  const int       n   = args.nIn[c];
  const int      *ids = args.inIds + args.inOff[c];
  double         *mp  = args.mat + args.matOff[c];
  const RbfParams rbf = args.rbf;
  double         *sx = dyn, *sy = dyn + n, *sz = dyn + 2 * n, *col = dyn + 3 * n;
  __shared__ int  bad;
  if (threadIdx.x == 0)
    bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double *p = args.inXyz + 3 * static_cast<long long>(ids[i]);
    sx[i]           = p[0];
    sy[i]           = p[1];
    sz[i]           = p[2];
  }
  __syncthreads();
  const double mz = args.dim == 3 ? 1.0 : 0.0;
  for (int j = 0; j < n; ++j) {
    const long long cs = colStart(j, n);
    for (int i = j + threadIdx.x; i < n; i += blockDim.x) {
      const double d2 = sqDist3Masked(sx[j], sy[j], sz[j], sx[i], sy[i], sz[i], 1.0, 1.0, mz);
      mp[cs + i - j]  = evalBasisT<KIND>(__dsqrt_rn(d2), rbf);
    }
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const long long ck = colStart(k, n);
    for (int i = k + threadIdx.x; i < n; i += blockDim.x)
      col[i] = mp[ck + i - k];
    __syncthreads();
    const double d = col[k];
    if (!(d > 0.0)) {
      if (threadIdx.x == 0)
        bad = 1;
      break;
    }
    const double rd = 1.0 / d;
    if (threadIdx.x == 0)
      mp[ck] = rd;
    // trailing update A(i,j) -= M(i,k) M(j,k) / d for j > k, i >= j
    const int       rem   = n - k - 1;
    const long long total = static_cast<long long>(rem) * (rem + 1) / 2;
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
      // map linear index e to (column jj, row offset) of the packed trailing triangle
      int       jj  = static_cast<int>((2.0 * rem + 1.0 - sqrt((2.0 * rem + 1.0) * (2.0 * rem + 1.0) - 8.0 * static_cast<double>(e))) * 0.5);
      long long cs0 = static_cast<long long>(jj) * rem - static_cast<long long>(jj) * (jj - 1) / 2;
      while (cs0 > e) {
        --jj;
        cs0 = static_cast<long long>(jj) * rem - static_cast<long long>(jj) * (jj - 1) / 2;
      }
      while (cs0 + (rem - jj) <= e) {
        cs0 += rem - jj;
        ++jj;
      }
      const int j = k + 1 + jj;
      const int i = j + static_cast<int>(e - cs0);
      double   *t = mp + colStart(j, n) + (i - j);
      *t          = fma(-col[i], col[j] * rd, *t);
    }
    __syncthreads();
  }
  __syncthreads();
  if (bad && threadIdx.x == 0)
    atomicMin(fail, c);
}

// =============================================================================================
// map kernels (shared pieces)
// =============================================================================================
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

// normalized partition-of-unity weight of out-vertex (x,y,z) in the cluster centred at (cx,cy,cz)
// (PartitionOfUnityMapping.hpp:321-336)
__device__ __forceinline__ double normalizedWeight(double x, double y, double z, double cx, double cy, double cz, double rinv, double wsum, int wcount)
{
  if (wsum <= 0.0)
    return __ddiv_rn(__ddiv_rn(1.0, static_cast<double>(wcount)), 1.0);
  const double w = pumWeight(__dsqrt_rn(sqDist3(cx, cy, cz, x, y, z)), rinv);
  return __ddiv_rn(w, wsum);
}

// Shared-memory layout of the map kernels (doubles):
//   mp   : n(n+1)/2   packed factor (only when it fits, STAGE_M)
//   sx,sy,sz : n each
//   b    : NC * n      right-hand sides / coefficients
//   red  : 16 * nWarps reduction scratch
//   (conservative) ox,oy,oz,ou[NC] : chunk of out-vertices, NT each
template <int KIND, int NC, int NT, bool STAGE_M>
__global__ void __launch_bounds__(NT) pumMapConsistentKernel(PumSolveArgs args, const int *__restrict__ list, const double *__restrict__ in, int dims, int c0,
                                                             double *__restrict__ out)
{
  extern __shared__ __align__(16) double dyn[];
  const int       c    = list[blockIdx.x];
  const int       n    = args.nIn[c];
  const int       m    = args.nOut[c];
  const int       tid  = threadIdx.x;
  const int      *ids  = args.inIds + args.inOff[c];
  const int      *oids = args.outIds + args.outOff[c];
  const double   *mg   = args.mat + args.matOff[c];
  const RbfParams rbf  = args.rbf;
  const PolyRec   rec  = args.poly[c];
  const double    cx = args.centers[3 * c], cy = args.centers[3 * c + 1], cz = args.centers[3 * c + 2];
  const double    rinv = 1.0 / args.radius;
  const long long tri  = static_cast<long long>(n) * (n + 1) / 2;

  double *ms  = dyn;
  double *sx  = STAGE_M ? dyn + tri : dyn;
  double *sy  = sx + n;
  double *sz  = sy + n;
  double *b   = sz + n;
  double *red = b + NC * n;

  if (m == 0)
    return; // nothing to evaluate (uniform per CTA)
  if (STAGE_M)
    for (long long e = tid; e < tri; e += NT)
      ms[e] = mg[e];
  for (int i = tid; i < n; i += NT) {
    const long long id = ids[i];
    const double   *p  = args.inXyz + 3 * id;
    sx[i]              = p[0];
    sy[i]              = p[1];
    sz[i]              = p[2];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      b[cc * n + i] = in[id * dims + c0 + cc]; // SphericalVertexCluster.hpp:312-320
  }
  __syncthreads();
  const double *mp = STAGE_M ? ms : mg;

  // ---- separated polynomial: beta = lstsq(Q, f); f <- f - Q beta (RadialBasisFctSolver.hpp:446-449) ----
  double beta[4][NC];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      beta[q][cc] = 0.0;
  if (rec.mode == 1) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) { // normal equations + one refinement step on the residual
      double t[4 * NC];
#pragma unroll
      for (int e = 0; e < 4 * NC; ++e)
        t[e] = 0.0;
      for (int i = tid; i < n; i += NT) {
        double q[4];
        polyRow(rec.mode, rec.mask, sx[i], sy[i], sz[i], cx, cy, cz, rinv, q);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double r = b[cc * n + i];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            r -= q[k] * beta[k][cc];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            t[k * NC + cc] += q[k] * r;
        }
      }
      blockReduceSum<4 * NC>(t, red, NT);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < 4; ++l)
            s += rec.G[4 * k + l] * t[l * NC + cc];
          beta[k][cc] += s;
        }
    }
  } else if (rec.mode == 2) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        double s = 0.0;
        for (int i = 0; i < n && i < 3; ++i)
          s += rec.G[4 * k + i] * b[cc * n + i];
        beta[k][cc] = s;
      }
  }
  if (rec.mode != 0) {
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      double q[4];
      polyRow(rec.mode, rec.mask, sx[i], sy[i], sz[i], cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        double r = b[cc * n + i];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          r -= q[k] * beta[k][cc];
        b[cc * n + i] = r;
      }
    }
    __syncthreads();
  }

  // ---- lambda = C^-1 f ----
  blockedSolve<NC>(mp, n, b, n, tid, NT);

  // ---- out = A lambda + V beta, weighted accumulation (SphericalVertexCluster.hpp:322-334) ----
  const double mz = args.dim == 3 ? 1.0 : 0.0;
  for (int o = tid; o < m; o += NT) {
    const long long gid = oids[o];
    const double   *p   = args.outXyz + 3 * gid;
    const double    x = p[0], y = p[1], z = p[2];
    double          acc[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      acc[cc] = 0.0;
    for (int j = 0; j < n; ++j) {
      const double d2  = sqDist3Masked(x, y, z, sx[j], sy[j], sz[j], 1.0, 1.0, mz);
      const double phi = evalBasisT<KIND>(__dsqrt_rn(d2), rbf);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
        acc[cc] = fma(phi, b[cc * n + j], acc[cc]);
    }
    if (rec.mode != 0) {
      double q[4];
      polyRow(rec.mode, rec.mask, x, y, z, cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          acc[cc] += q[k] * beta[k][cc];
    }
    const double w = normalizedWeight(x, y, z, cx, cy, cz, rinv, args.wsum[gid], args.wcount[gid]);
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      atomicAdd(&out[gid * dims + c0 + cc], acc[cc] * w);
  }
}

template <int KIND, int NC, int NT, bool STAGE_M>
__global__ void __launch_bounds__(NT) pumMapConservativeKernel(PumSolveArgs args, const int *__restrict__ list, const double *__restrict__ in, int dims, int c0,
                                                               double *__restrict__ out)
{
  extern __shared__ __align__(16) double dyn[];
  const int       c    = list[blockIdx.x];
  const int       n    = args.nIn[c];
  const int       m    = args.nOut[c];
  const int       tid  = threadIdx.x;
  const int      *ids  = args.inIds + args.inOff[c];
  const int      *oids = args.outIds + args.outOff[c];
  const double   *mg   = args.mat + args.matOff[c];
  const RbfParams rbf  = args.rbf;
  const PolyRec   rec  = args.poly[c];
  const double    cx = args.centers[3 * c], cy = args.centers[3 * c + 1], cz = args.centers[3 * c + 2];
  const double    rinv = 1.0 / args.radius;
  const long long tri  = static_cast<long long>(n) * (n + 1) / 2;

  double *ms  = dyn;
  double *sx  = STAGE_M ? dyn + tri : dyn;
  double *sy  = sx + n;
  double *sz  = sy + n;
  double *b   = sz + n;
  double *red = b + NC * n;
  double *ox  = red + 16 * (NT / 32);
  double *oy  = ox + NT;
  double *oz  = oy + NT;
  double *ou  = oz + NT; // NC * NT

  if (m == 0)
    return;
  if (STAGE_M)
    for (long long e = tid; e < tri; e += NT)
      ms[e] = mg[e];
  for (int i = tid; i < n; i += NT) {
    const double *p = args.inXyz + 3 * static_cast<long long>(ids[i]);
    sx[i]           = p[0];
    sy[i]           = p[1];
    sz[i]           = p[2];
  }
  const double *mp = STAGE_M ? ms : mg;
  const double  mz = args.dim == 3 ? 1.0 : 0.0;

  // ---- Au = A^T (w .* u), epsilon = V^T (w .* u) (RadialBasisFctSolver.hpp:420, :428) ----
  double eps[4 * NC];
#pragma unroll
  for (int e = 0; e < 4 * NC; ++e)
    eps[e] = 0.0;
  const int rowsPerThread = (n + NT - 1) / NT; // 1 for the register-tiled classes
  for (int i = tid; i < n; i += NT)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      b[cc * n + i] = 0.0;
  for (int o0 = 0; o0 < m; o0 += NT) {
    __syncthreads();
    const int o = o0 + tid;
    if (o < m) {
      const long long gid = oids[o];
      const double   *p   = args.outXyz + 3 * gid;
      const double    x = p[0], y = p[1], z = p[2];
      ox[tid] = x;
      oy[tid] = y;
      oz[tid] = z;
      // SphericalVertexCluster.hpp:219-227: in(i,c) = data * normalizedWeight
      const double w = normalizedWeight(x, y, z, cx, cy, cz, rinv, args.wsum[gid], args.wcount[gid]);
      double       q[4];
      if (rec.mode != 0)
        polyRow(rec.mode, rec.mask, x, y, z, cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        const double u   = in[gid * dims + c0 + cc] * w;
        ou[cc * NT + tid] = u;
        if (rec.mode != 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            eps[k * NC + cc] += q[k] * u;
        }
      }
    }
    __syncthreads();
    const int chunk = min(NT, m - o0);
    for (int r = 0; r < rowsPerThread; ++r) {
      const int j = tid + r * NT;
      if (j < n) {
        double acc[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          acc[cc] = b[cc * n + j];
        const double xj = sx[j], yj = sy[j], zj = sz[j];
        for (int oo = 0; oo < chunk; ++oo) {
          const double d2  = sqDist3Masked(ox[oo], oy[oo], oz[oo], xj, yj, zj, 1.0, 1.0, mz);
          const double phi = evalBasisT<KIND>(__dsqrt_rn(d2), rbf);
#pragma unroll
          for (int cc = 0; cc < NC; ++cc)
            acc[cc] = fma(phi, ou[cc * NT + oo], acc[cc]);
        }
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          b[cc * n + j] = acc[cc];
      }
    }
  }
  if (rec.mode != 0)
    blockReduceSum<4 * NC>(eps, red, NT);
  __syncthreads();

  // ---- mu = C^-1 Au ----
  blockedSolve<NC>(mp, n, b, n, tid, NT);

  // ---- out = mu + Q (Q^T Q)^-1 (epsilon - Q^T mu)  (RadialBasisFctSolver.hpp:427-436) ----
  double gamma[4][NC];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      gamma[k][cc] = 0.0;
  if (rec.mode == 1) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      // tau = epsilon - Q^T (mu + Q gamma)
      double t[4 * NC];
#pragma unroll
      for (int e = 0; e < 4 * NC; ++e)
        t[e] = 0.0;
      for (int i = tid; i < n; i += NT) {
        double q[4];
        polyRow(rec.mode, rec.mask, sx[i], sy[i], sz[i], cx, cy, cz, rinv, q);
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double v = b[cc * n + i];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            v += q[k] * gamma[k][cc];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            t[k * NC + cc] += q[k] * v;
        }
      }
      blockReduceSum<4 * NC>(t, red, NT);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < 4; ++l)
            s += rec.G[4 * k + l] * (eps[l * NC + cc] - t[l * NC + cc]);
          gamma[k][cc] += s;
        }
    }
  }
  for (int i = tid; i < n; i += NT) {
    const long long gid = ids[i];
    double          res[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      res[cc] = b[cc * n + i];
    if (rec.mode == 1) {
      double q[4];
      polyRow(rec.mode, rec.mask, sx[i], sy[i], sz[i], cx, cy, cz, rinv, q);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          res[cc] += q[k] * gamma[k][cc];
    } else if (rec.mode == 2) {
      // y = S^T (epsilon - Q^T mu) with the raw basis; n <= 3 rows
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        double tau[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tau[k] = eps[k * NC + cc];
        for (int r = 0; r < n && r < 3; ++r) {
          double q[4];
          polyRow(rec.mode, rec.mask, sx[r], sy[r], sz[r], cx, cy, cz, rinv, q);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tau[k] -= q[k] * b[cc * n + r];
        }
        double yv = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          yv += rec.G[4 * k + i] * tau[k];
        res[cc] += yv;
      }
    }
    // SphericalVertexCluster.hpp:233-239
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
      atomicAdd(&out[gid * dims + c0 + cc], res[cc]);
  }
}

// =============================================================================================
// host launchers
// =============================================================================================
template <typename K>
void allowSmem(K kernel, size_t bytes)
{
  if (bytes > 48 * 1024)
    RBF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
}

constexpr size_t kMaxSmem = 200 * 1024;

template <int KIND, int NC>
void launchConsistent(const PumSolveArgs &a, const int *list, int64_t count, int sizeClass, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
{
  const size_t tri = static_cast<size_t>(maxN) * (maxN + 1) / 2;
  if (sizeClass == kClassSmall) {
    constexpr int NT = 64;
    const size_t  sm = (tri + (3 + NC) * static_cast<size_t>(maxN) + 16 * (NT / 32)) * sizeof(double);
    allowSmem(pumMapConsistentKernel<KIND, NC, NT, true>, sm);
    pumMapConsistentKernel<KIND, NC, NT, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else if (sizeClass == kClassMedium) {
    constexpr int NT = 128;
    const size_t  sm = (tri + (3 + NC) * static_cast<size_t>(maxN) + 16 * (NT / 32)) * sizeof(double);
    allowSmem(pumMapConsistentKernel<KIND, NC, NT, true>, sm);
    pumMapConsistentKernel<KIND, NC, NT, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else {
    constexpr int NT = 256;
    const size_t  sm = ((3 + NC) * static_cast<size_t>(maxN) + 16 * (NT / 32)) * sizeof(double);
    if (sm > kMaxSmem)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size.");
    allowSmem(pumMapConsistentKernel<KIND, NC, NT, false>, sm);
    pumMapConsistentKernel<KIND, NC, NT, false><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  }
  RBF_LAUNCHED();
}

template <int KIND, int NC>
void launchConservative(const PumSolveArgs &a, const int *list, int64_t count, int sizeClass, int maxN, const double *in, int dims, int c0, double *out, cudaStream_t s)
{
  const size_t tri = static_cast<size_t>(maxN) * (maxN + 1) / 2;
  if (sizeClass == kClassSmall) {
    constexpr int NT = 64;
    const size_t  sm = (tri + (3 + NC) * static_cast<size_t>(maxN) + 16 * (NT / 32) + (3 + NC) * NT) * sizeof(double);
    allowSmem(pumMapConservativeKernel<KIND, NC, NT, true>, sm);
    pumMapConservativeKernel<KIND, NC, NT, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else if (sizeClass == kClassMedium) {
    constexpr int NT = 128;
    const size_t  sm = (tri + (3 + NC) * static_cast<size_t>(maxN) + 16 * (NT / 32) + (3 + NC) * NT) * sizeof(double);
    allowSmem(pumMapConservativeKernel<KIND, NC, NT, true>, sm);
    pumMapConservativeKernel<KIND, NC, NT, true><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  } else {
    constexpr int NT = 256;
    const size_t  sm = ((3 + NC) * static_cast<size_t>(maxN) + 16 * (NT / 32) + (3 + NC) * NT) * sizeof(double);
    if (sm > kMaxSmem)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size.");
    allowSmem(pumMapConservativeKernel<KIND, NC, NT, false>, sm);
    pumMapConservativeKernel<KIND, NC, NT, false><<<static_cast<unsigned>(count), NT, sm, s>>>(a, list, in, dims, c0, out);
  }
  RBF_LAUNCHED();
}

template <int KIND>
void mapDispatch(bool conservative, const PumSolveArgs &a, const int *list, int64_t count, int sizeClass, int maxN, const double *in, int dims, double *out, cudaStream_t s)
{
  // components are processed in groups of up to 3 per launch
  for (int c0 = 0; c0 < dims;) {
    const int nc = std::min(3, dims - c0);
    if (conservative) {
      if (nc == 1) launchConservative<KIND, 1>(a, list, count, sizeClass, maxN, in, dims, c0, out, s);
      else if (nc == 2) launchConservative<KIND, 2>(a, list, count, sizeClass, maxN, in, dims, c0, out, s);
      else launchConservative<KIND, 3>(a, list, count, sizeClass, maxN, in, dims, c0, out, s);
    } else {
      if (nc == 1) launchConsistent<KIND, 1>(a, list, count, sizeClass, maxN, in, dims, c0, out, s);
      else if (nc == 2) launchConsistent<KIND, 2>(a, list, count, sizeClass, maxN, in, dims, c0, out, s);
      else launchConsistent<KIND, 3>(a, list, count, sizeClass, maxN, in, dims, c0, out, s);
    }
    c0 += nc;
  }
}

} // namespace

void pumPolySetup(const PumSolveArgs &a, int64_t nClusters, int polynomial, cudaStream_t s)
{
  if (nClusters <= 0)
    return;
  if (polynomial == RBFMAP_POLYNOMIAL_SEPARATE)
    polySetupKernel<<<divUp(nClusters, 128), 128, 0, s>>>(a, nClusters);
  else
    polyOffKernel<<<divUp(nClusters, 256), 256, 0, s>>>(a.poly, nClusters);
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

void pumFactor(const PumSolveArgs &a, const int *list, int64_t count, int sizeClass, int maxN, int *fail, cudaStream_t s)
{
  if (count <= 0)
    return;
  RBF_DISPATCH_BASIS(a.rbf.kind, {
    if (sizeClass == kClassSmall) {
      pumFactorKernel<KIND, 8><<<static_cast<unsigned>(count), 64, 0, s>>>(a, list, fail);
    } else if (sizeClass == kClassMedium) {
      pumFactorKernel<KIND, 16><<<static_cast<unsigned>(count), 256, 0, s>>>(a, list, fail);
    } else {
      const size_t sm = 4 * static_cast<size_t>(maxN) * sizeof(double);
      if (sm > kMaxSmem)
        throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size.");
      allowSmem(pumFactorGenericKernel<KIND>, sm);
      pumFactorGenericKernel<KIND><<<static_cast<unsigned>(count), 256, sm, s>>>(a, list, fail);
    }
  });
  RBF_LAUNCHED();
  RBF_CUDA(cudaGetLastError());
}

void pumMapConsistent(const PumSolveArgs &a, const int *list, int64_t count, int sizeClass, int maxN, const double *in, int dims, double *out, cudaStream_t s)
{
  if (count <= 0)
    return;
  RBF_DISPATCH_BASIS(a.rbf.kind, mapDispatch<KIND>(false, a, list, count, sizeClass, maxN, in, dims, out, s));
  RBF_CUDA(cudaGetLastError());
}

void pumMapConservative(const PumSolveArgs &a, const int *list, int64_t count, int sizeClass, int maxN, const double *in, int dims, double *out, cudaStream_t s)
{
  if (count <= 0)
    return;
  RBF_DISPATCH_BASIS(a.rbf.kind, mapDispatch<KIND>(true, a, list, count, sizeClass, maxN, in, dims, out, s));
  RBF_CUDA(cudaGetLastError());
}

} // namespace rbfmap
