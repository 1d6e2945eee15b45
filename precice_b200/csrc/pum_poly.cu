// pum_poly.cu — per-cluster set-up of the separated polynomial (one thread per cluster).
//
// Replaces RadialBasisFctSolver.hpp:377-410: polynomial matrix Q = [1, x_active...], condition number
// sigma_max / max(sigma_min, 1e-14) > 1e5 -> drop the active axis with the smallest extent (reduceActiveAxis, :115-137),
// then the decomposition used by solveConsistent / solveConservative (:409, :427-436, :446-449).
// (Kokkos prior art: do_batched_qr, device/KokkosPUMKernels_Impl.hpp:177-276.)
#include "pum_device.cuh"
#include "small_dense.cuh"

#include <algorithm>
#include <climits>

namespace rbfmap {
namespace {
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
  // ... and of the centred/scaled basis [1, (x_a - c_a)/r] over all three axes: the Gram matrix of the axes that survive
  // the condition test below is a sub-matrix of it, so the vertices are gathered only once
  const double cx = a.centers[3 * c], cy = a.centers[3 * c + 1], cz = a.centers[3 * c + 2];
  const double rinv = 1.0 / a.radius;
  double       mc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      mc[i][j] = 0.0;
  // The vertices are accumulated strictly in order (the sums are bit-identical to the one-at-a-time loop), but their
  // coordinates are fetched four vertices ahead: one thread walks its own 1.2 KB record, and with one dependent
  // load -> accumulate round trip per vertex the kernel ran at the L2 latency (0.83 ms at 10M vertices).
  const double *rec0 = a.inRec + 3 * a.inOff[c];
  auto          accumulate = [&](int i, double x0, double x1, double x2) {
    const double x[3]  = {x0, x1, x2};
    const double q[4]  = {1.0, x0, x1, x2};
    const double qc[4] = {1.0, (x0 - cx) * rinv, (x1 - cy) * rinv, (x2 - cz) * rinv};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s2 = r; s2 < 4; ++s2)
        mc[r][s2] += qc[r] * qc[s2];
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
  };
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    double v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k)
      v[k] = __ldg(rec0 + 3 * i + k);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      accumulate(i + k, v[3 * k], v[3 * k + 1], v[3 * k + 2]);
  }
  for (; i < n; ++i)
    accumulate(i, __ldg(rec0 + 3 * i), __ldg(rec0 + 3 * i + 1), __ldg(rec0 + 3 * i + 2));

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
    // Gram matrix of the centred/scaled basis over the active axes (polyRow packs the active axes in increasing
    // order into q[1..p-1]), then its inverse
    int sel[4] = {0, 0, 0, 0}; // basis index -> slot of the full basis
    {
      int k = 1;
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (mask & (1 << d)) {
          if (k == 1) sel[1] = d + 1;
          else if (k == 2) sel[2] = d + 1;
          else sel[3] = d + 1;
          ++k;
        }
    }
    double g[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) {
        double v = 0.0;
        if (r < p && s2 < p) {
          const int fr = sel[r], fs = sel[s2];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w)
              if (u == (fr < fs ? fr : fs) && w == (fr < fs ? fs : fr))
                v = mc[u][w];
        }
        g[r][s2] = v;
      }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r >= p)
        g[r][r] = 1.0;
    double normG = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      normG = fmax(normG, fabs(g[r][0]) + fabs(g[r][1]) + fabs(g[r][2]) + fabs(g[r][3]));
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
double normI = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      normI = fmax(normI, fabs(inv[r][0]) + fabs(inv[r][1]) + fabs(inv[r][2]) + fabs(inv[r][3]));
    // the map kernels add one refinement step to the normal-equation solve when the centred Gram matrix is not
    // well conditioned (nearly degenerate clusters that just passed the 1e5 test above)
    rec.pad = (normG * normI > 1.0e3) ? 1 : 0;
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


__global__ void distSqrtKernel(const double *__restrict__ x, long long n, double *__restrict__ out)
{
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n)
    out[i] = distSqrt(x[i]);
}
} // namespace

// Test hook: the pair evaluation the per-cluster algebra kernels use (pairBasis / pairBasisBatch of pum_device.cuh: fused
// multiply-adds, reciprocal-root seed, integer clamp at the support), NOT the operation-by-operation evalBasis.
template <int KIND, bool DIM3>
__global__ void pairBasisDebugKernel(RbfParams f, const double *__restrict__ a, const double *__restrict__ b, long long n, int batched, double *__restrict__ out)
{
  const double    invR2 = f.p1 * f.p1;
  const long long t     = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (!batched) {
    if (t < n)
      out[t] = pairBasis<KIND, DIM3>(a[3 * t], a[3 * t + 1], a[3 * t + 2], b[3 * t], b[3 * t + 1], b[3 * t + 2], f, invR2);
    return;
  }
  // four consecutive pairs per thread through the staged batch form (what the evaluation loops of the map kernels call)
  const long long i0 = 4 * t;
  if (i0 >= n)
    return;
  double ax[4], ay[4], az[4], bx[4], by[4], bz[4], phi[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long i = min(i0 + u, n - 1);
    ax[u] = a[3 * i], ay[u] = a[3 * i + 1], az[u] = a[3 * i + 2];
    bx[u] = b[3 * i], by[u] = b[3 * i + 1], bz[u] = b[3 * i + 2];
  }
  pairBasisBatch<KIND, DIM3, 4>(ax, ay, az, bx, by, bz, f, invR2, phi);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (i0 + u < n)
      out[i0 + u] = phi[u];
}

void debugPairBasis(const RbfParams &f, bool dim3, const double *dA, const double *dB, long long n, int batched, double *dOut)
{
  const int grid = divUp(batched ? (n + 3) / 4 : n, 256);
  RBF_DISPATCH_HOT_BASIS(f.kind, {
    if (dim3)
      pairBasisDebugKernel<KIND, true><<<grid, 256>>>(f, dA, dB, n, batched, dOut);
    else
      pairBasisDebugKernel<KIND, false><<<grid, 256>>>(f, dA, dB, n, batched, dOut);
  });
  RBF_CUDA(cudaGetLastError());
}

void debugDistSqrt(const double *dX, long long n, double *dOut)
{
  distSqrtKernel<<<divUp(n, 256), 256>>>(dX, n, dOut);
  RBF_CUDA(cudaGetLastError());
}

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

} // namespace rbfmap
