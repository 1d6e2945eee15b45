// rbf_oracle.hpp — CPU ORACLE (test infrastructure, NOT the product).
//
// A from-scratch, dependency-free C++17 restatement of the algorithm of preCICE's
// radial-basis-function data-mapping hot path (preCICE 3.4.1, /root/reference). It exists only
// to *check* the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
// --impl reference leg). The shipped library (precice_b200/) never includes, links or calls it.
//
// The reference cannot be compiled in this image (needs Eigen, Boost, libxml2: see DESIGN.md),
// so every Eigen/Boost facility it uses is restated here:
//   Eigen::LLT                  -> orc::Cholesky            (RadialBasisFctSolver.hpp:27,352)
//   Eigen::ColPivHouseholderQR  -> orc::ColPivQR            (RadialBasisFctSolver.hpp:27,356,409)
//   Eigen::JacobiSVD (cond.)    -> orc::singularValues      (RadialBasisFctSolver.hpp:389-392)
//   Boost.Geometry R*-tree      -> orc::PointGrid           (query/Index.cpp:171-253, 369-391)
//   boost::flat_set             -> sorted std::vector<int>  (SphericalVertexCluster.hpp:115-118,161-162)
// Parity is PINNED by the reference's own golden literals (tests/test_oracle_golden.py):
// PartitionOfUnityMappingTest.cpp:375-377,475,617,1191,1308,1774, PartitionOfUnityClusteringTest.cpp:51-94,
// RadialBasisFctMappingTest.cpp:1461-1473,1525-1546 and the linear-reproduction tables.
//
// Conventions: coordinates are double[N][3] (2-D meshes carry z = 0, mesh/Vertex.hpp:75-77);
// values are vertex-major interleaved, values[i*dims + c] (RadialBasisFctMapping.hpp:257-261).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace orc {

// math/differences.hpp:8
constexpr double NUMERICAL_ZERO_DIFFERENCE = 1.0e-14;

// Same numeric values as mapping/config/MappingConfigurationTypes.hpp:11-30 and Mapping.hpp:30-35
enum class Polynomial : int { ON = 0, OFF = 1, SEPARATE = 2 };
enum class BasisFunction : int {
  WendlandC0 = 0,
  WendlandC2,
  WendlandC4,
  WendlandC6,
  WendlandC8,
  ThinPlateSplines,
  Multiquadrics,
  InverseMultiquadrics,
  VolumeSplines,
  Gaussian,
  CompactThinPlateSplinesC2
};
enum class Constraint : int { CONSISTENT = 0, CONSERVATIVE = 1, SCALED_CONSISTENT_SURFACE = 2, SCALED_CONSISTENT_VOLUME = 3 };

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------------------------------------
// math/math.hpp:21-40 — integer power by repeated squaring (same multiplication order)
template <int iexp>
inline double pow_int(double base)
{
  if (iexp == 0)
    return 1.0;
  int    e = iexp;
  double x = base, y = 1.0;
  while (e > 1) {
    if (e % 2 == 1)
      y *= x;
    x *= x;
    e /= 2;
  }
  return x * y;
}

// ---------------------------------------------------------------------------------------------
// impl/BasisFunctions.hpp:40-44 (RadialBasisParameters) and :87-607 (functors)
struct Rbf {
  BasisFunction kind = BasisFunction::WendlandC0;
  double        p1 = 0, p2 = 0, p3 = 0;

  // Constructor semantics of each functor: `a` is the support radius for compact functions,
  // the shape parameter for (inverse) multiquadrics and the Gaussian.
  static Rbf make(BasisFunction k, double a, double gaussSupport = std::numeric_limits<double>::infinity())
  {
    Rbf f;
    f.kind = k;
    switch (k) {
    case BasisFunction::ThinPlateSplines:
    case BasisFunction::VolumeSplines:
      break;
    case BasisFunction::Multiquadrics: // :120-124
      f.p1 = pow_int<2>(a);
      break;
    case BasisFunction::InverseMultiquadrics: // :159-166
      if (!(a > 0.0))
        throw std::invalid_argument("Shape parameter for radial-basis-function inverse multiquadric has to be larger than zero. Please update the \"shape-parameter\" attribute.");
      f.p1 = pow_int<2>(a);
      break;
    case BasisFunction::Gaussian: { // :229-252
      if (!(a > 0.0))
        throw std::invalid_argument("Shape parameter for radial-basis-function gaussian has to be larger than zero. Please update the \"shape-parameter\" attribute.");
      if (!(gaussSupport > 0.0))
        throw std::invalid_argument("Support radius for radial-basis-function gaussian has to be larger than zero. Please update the \"support-radius\" attribute.");
      const double threshold = std::sqrt(-std::log(1e-9)) / a;
      f.p1                   = a;
      f.p2                   = std::min(gaussSupport, threshold);
      f.p3                   = 0.0;
      if (gaussSupport < std::numeric_limits<double>::infinity()) {
        f.p3 = f(gaussSupport); // deltaY evaluated with deltaY = 0
      }
      break;
    }
    default: // compact polynomials and compact TPS: r_inv = 1/supportRadius
      if (!(a > 0.0))
        throw std::invalid_argument("Support radius for the compact radial-basis-function has to be larger than zero. Please update the \"support-radius\" attribute.");
      f.p1 = 1.0 / a;
    }
    return f;
  }

  bool isStrictlyPositiveDefinite() const
  {
    switch (kind) {
    case BasisFunction::ThinPlateSplines:
    case BasisFunction::Multiquadrics:
    case BasisFunction::VolumeSplines:
      return false;
    default:
      return true;
    }
  }

  bool hasCompactSupport() const
  {
    switch (kind) {
    case BasisFunction::ThinPlateSplines:
    case BasisFunction::Multiquadrics:
    case BasisFunction::InverseMultiquadrics:
    case BasisFunction::VolumeSplines:
      return false;
    default:
      return true;
    }
  }

  double getSupportRadius() const
  {
    if (!hasCompactSupport())
      return std::numeric_limits<double>::max();
    if (kind == BasisFunction::Gaussian)
      return p2;
    return 1.0 / p1;
  }

  // Same floating-point operation sequence as the reference functors (no contraction:
  // the oracle is compiled with -ffp-contract=off; std::fma only where the reference has one).
  double operator()(double radius) const
  {
    switch (kind) {
    case BasisFunction::ThinPlateSplines: // :95-99
      return std::log(std::max(radius, 1.0e-14)) * pow_int<2>(radius);
    case BasisFunction::Multiquadrics: // :131-135
      return std::sqrt(p1 + pow_int<2>(radius));
    case BasisFunction::InverseMultiquadrics: // :172-176
      return 1.0 / std::sqrt(p1 + pow_int<2>(radius));
    case BasisFunction::VolumeSplines: // :205-208
      return std::abs(radius);
    case BasisFunction::Gaussian: // :263-272
      if (radius > p2)
        return 0.0;
      return std::exp(-pow_int<2>(p1 * radius)) - p3;
    case BasisFunction::CompactThinPlateSplinesC2: { // :327-335
      const double p = radius * p1;
      if (p >= 1)
        return 0.0;
      return 1.0 - 30.0 * pow_int<2>(p) - 10.0 * pow_int<3>(p) + 45.0 * pow_int<4>(p) - 6.0 * pow_int<5>(p) - pow_int<3>(p) * 60.0 * std::log(std::max(p, 1.0e-14));
    }
    case BasisFunction::WendlandC0: { // :381-388
      const double p = radius * p1;
      if (p >= 1)
        return 0.0;
      return pow_int<2>(1.0 - p);
    }
    case BasisFunction::WendlandC2: { // :434-441
      const double p = radius * p1;
      if (p >= 1)
        return 0.0;
      return pow_int<4>(1.0 - p) * std::fma(4.0, p, 1.0);
    }
    case BasisFunction::WendlandC4: { // :487-494
      const double p = radius * p1;
      if (p >= 1)
        return 0.0;
      return pow_int<6>(1.0 - p) * (35 * pow_int<2>(p) + std::fma(18.0, p, 3.0));
    }
    case BasisFunction::WendlandC6: { // :538-545
      const double p = radius * p1;
      if (p >= 1)
        return 0.0;
      return pow_int<8>(1.0 - p) * (32.0 * pow_int<3>(p) + 25.0 * pow_int<2>(p) + std::fma(8.0, p, 1.0));
    }
    case BasisFunction::WendlandC8: { // :591-598
      const double p = radius * p1;
      if (p >= 1)
        return 0.0;
      return pow_int<10>(1.0 - p) * (1287.0 * pow_int<4>(p) + 1350.0 * pow_int<3>(p) + 630.0 * pow_int<2>(p) + 150.0 * p + 15);
    }
    }
    return 0.0;
  }
};

// ---------------------------------------------------------------------------------------------
// RadialBasisFctSolver.hpp:102-113
using Vec3 = std::array<double, 3>;
inline double squaredDifference(const double *u, const double *v, const std::array<bool, 3> &active = {{true, true, true}})
{
  double w[3];
  for (int d = 0; d < 3; ++d)
    w[d] = (u[d] - v[d]) * static_cast<int>(active[d]);
  double s = 0.0; // std::inner_product order
  for (int d = 0; d < 3; ++d)
    s = s + w[d] * w[d];
  return s;
}

// ---------------------------------------------------------------------------------------------
// Dense column-major matrix (stands in for Eigen::MatrixXd)
struct Mat {
  int                 r = 0, c = 0;
  std::vector<double> a;
  Mat() = default;
  Mat(int rows, int cols)
      : r(rows), c(cols), a(static_cast<size_t>(rows) * cols, 0.0) {}
  double       &operator()(int i, int j) { return a[static_cast<size_t>(j) * r + i]; }
  const double &operator()(int i, int j) const { return a[static_cast<size_t>(j) * r + i]; }
  double       *col(int j) { return a.data() + static_cast<size_t>(j) * r; }
  const double *col(int j) const { return a.data() + static_cast<size_t>(j) * r; }
  size_t        size() const { return a.size(); }
};

// out = A * B
inline Mat matmul(const Mat &A, const Mat &B)
{
  Mat C(A.r, B.c);
  for (int j = 0; j < B.c; ++j)
    for (int k = 0; k < A.c; ++k) {
      const double  b  = B(k, j);
      const double *ak = A.col(k);
      double       *cj = C.col(j);
      for (int i = 0; i < A.r; ++i)
        cj[i] += ak[i] * b;
    }
  return C;
}

// out = A^T * B
inline Mat matmulT(const Mat &A, const Mat &B)
{
  Mat C(A.c, B.c);
  for (int j = 0; j < B.c; ++j)
    for (int i = 0; i < A.c; ++i) {
      const double *ai = A.col(i);
      const double *bj = B.col(j);
      double        s  = 0.0;
      for (int k = 0; k < A.r; ++k)
        s += ai[k] * bj[k];
      C(i, j) = s;
    }
  return C;
}

// ---------------------------------------------------------------------------------------------
// Lower Cholesky A = L L^T (stands in for Eigen::LLT; used at RadialBasisFctSolver.hpp:352-353).
// Column-oriented left-looking sweep; `ok` mirrors `info() == Success`.
struct Cholesky {
  Mat  L;
  bool ok = false;

  void compute(Mat A)
  {
    const int n = A.r;
    ok          = true;
    // blocked right-looking variant: panel width nb, keeps the trailing update cache friendly
    const int nb = 48;
    for (int k0 = 0; k0 < n; k0 += nb) {
      const int kb = std::min(nb, n - k0);
      // factor the diagonal block + panel below it (unblocked, column by column)
      for (int k = k0; k < k0 + kb; ++k) {
        double *ck = A.col(k);
        // subtract contributions of the columns of this panel left of k
        for (int j = k0; j < k; ++j) {
          const double *cj  = A.col(j);
          const double  ljk = cj[k];
          if (ljk != 0.0)
            for (int i = k; i < n; ++i)
              ck[i] -= cj[i] * ljk;
        }
        double x = ck[k];
        if (!(x > 0.0)) {
          ok = false;
          L  = std::move(A);
          return;
        }
        x     = std::sqrt(x);
        ck[k] = x;
        for (int i = k + 1; i < n; ++i)
          ck[i] /= x;
      }
      // trailing update: A22 -= L21 L21^T (lower part only)
      const int t0 = k0 + kb;
      for (int j = t0; j < n; ++j) {
        double *cj = A.col(j);
        for (int k = k0; k < k0 + kb; ++k) {
          const double *ck  = A.col(k);
          const double  ljk = ck[j];
          if (ljk != 0.0)
            for (int i = j; i < n; ++i)
              cj[i] -= ck[i] * ljk;
        }
      }
    }
    L = std::move(A);
  }

  // Solves (L L^T) X = B
  Mat solve(const Mat &B) const
  {
    const int n = L.r;
    Mat       X = B;
    for (int c = 0; c < X.c; ++c) {
      double *x = X.col(c);
      for (int k = 0; k < n; ++k) { // L y = b (column sweep)
        const double *lk = L.col(k);
        x[k] /= lk[k];
        const double xk = x[k];
        for (int i = k + 1; i < n; ++i)
          x[i] -= lk[i] * xk;
      }
      for (int k = n - 1; k >= 0; --k) { // L^T x = y (dot products with column k)
        const double *lk = L.col(k);
        double        s  = x[k];
        for (int i = k + 1; i < n; ++i)
          s -= lk[i] * x[i];
        x[k] = s / lk[k];
      }
    }
    return X;
  }
};

// ---------------------------------------------------------------------------------------------
// Householder QR with column pivoting, restating the published algorithm of
// Eigen::ColPivHouseholderQR (Eigen 3.4, third-party, not under /root/reference): greedy pivot
// on the largest remaining column norm with norm down-dating, reflector H = I - tau v v^T with
// v(0) = 1. solve() gives the basic least-squares solution on the leading `nonzeroPivots`
// columns; solveTransposed() the minimum-norm solution of A^T x = b.
// Call sites: RadialBasisFctSolver.hpp:356-357 (C for non-SPD functions), :409 (Q), :435, :448.
struct ColPivQR {
  Mat                 qr;
  std::vector<double> tau;
  std::vector<int>    perm; // perm[i] = original column placed at position i
  int                 nonzeroPivots = 0;
  double              maxPivot      = 0.0;

  void compute(Mat A)
  {
    const int rows = A.r, cols = A.c, size = std::min(rows, cols);
    tau.assign(size, 0.0);
    perm.resize(cols);
    std::iota(perm.begin(), perm.end(), 0);
    std::vector<double> normUpd(cols), normDir(cols);
    for (int k = 0; k < cols; ++k) {
      double s = 0;
      for (int i = 0; i < rows; ++i)
        s += A(i, k) * A(i, k);
      normUpd[k] = normDir[k] = std::sqrt(s);
    }
    const double eps             = std::numeric_limits<double>::epsilon();
    const double maxNorm         = cols ? *std::max_element(normUpd.begin(), normUpd.end()) : 0.0;
    const double thresholdHelper = (maxNorm * eps) * (maxNorm * eps) / static_cast<double>(rows);
    const double downdateThresh  = std::sqrt(eps);
    nonzeroPivots                = size;
    maxPivot                     = 0.0;
    std::vector<double> tmp(cols);
    for (int k = 0; k < size; ++k) {
      int    big  = k;
      double best = normUpd[k];
      for (int j = k + 1; j < cols; ++j)
        if (normUpd[j] > best) {
          best = normUpd[j];
          big  = j;
        }
      const double bestSq = best * best;
      if (nonzeroPivots == size && bestSq < thresholdHelper * static_cast<double>(rows - k))
        nonzeroPivots = k;
      if (big != k) {
        for (int i = 0; i < rows; ++i)
          std::swap(A(i, k), A(i, big));
        std::swap(normUpd[k], normUpd[big]);
        std::swap(normDir[k], normDir[big]);
        std::swap(perm[k], perm[big]);
      }
      // Householder reflector for column k, rows k..rows-1
      double tailSq = 0;
      for (int i = k + 1; i < rows; ++i)
        tailSq += A(i, k) * A(i, k);
      const double c0 = A(k, k);
      double       beta, t;
      if (tailSq <= std::numeric_limits<double>::min()) {
        t    = 0.0;
        beta = c0;
        for (int i = k + 1; i < rows; ++i)
          A(i, k) = 0.0;
      } else {
        beta = std::sqrt(c0 * c0 + tailSq);
        if (c0 >= 0)
          beta = -beta;
        for (int i = k + 1; i < rows; ++i)
          A(i, k) /= (c0 - beta);
        t = (beta - c0) / beta;
      }
      tau[k]  = t;
      A(k, k) = beta;
      if (std::abs(beta) > maxPivot)
        maxPivot = std::abs(beta);
      // apply H to the remaining columns
      if (t != 0.0) {
        for (int j = k + 1; j < cols; ++j) {
          double s = A(k, j);
          for (int i = k + 1; i < rows; ++i)
            s += A(i, k) * A(i, j);
          s *= t;
          A(k, j) -= s;
          for (int i = k + 1; i < rows; ++i)
            A(i, j) -= s * A(i, k);
        }
      }
      // norm down-dating
      for (int j = k + 1; j < cols; ++j) {
        if (normUpd[j] != 0.0) {
          double temp = std::abs(A(k, j)) / normUpd[j];
          temp        = (1.0 + temp) * (1.0 - temp);
          temp        = temp < 0.0 ? 0.0 : temp;
          const double r2    = normUpd[j] / normDir[j];
          const double temp2 = temp * r2 * r2;
          if (temp2 <= downdateThresh) {
            double s = 0;
            for (int i = k + 1; i < rows; ++i)
              s += A(i, j) * A(i, j);
            normDir[j] = std::sqrt(s);
            normUpd[j] = normDir[j];
          } else {
            normUpd[j] *= std::sqrt(temp);
          }
        }
      }
    }
    qr = std::move(A);
  }

  int rank() const
  {
    const int    size      = std::min(qr.r, qr.c);
    const double threshold = std::numeric_limits<double>::epsilon() * static_cast<double>(size);
    const double pre       = std::abs(maxPivot) * threshold;
    int          res       = 0;
    for (int i = 0; i < nonzeroPivots; ++i)
      res += (std::abs(qr(i, i)) > pre) ? 1 : 0;
    return res;
  }

  bool isInvertible() const { return rank() == qr.c && rank() == qr.r; }

  // c := Q^T c  (first `len` reflectors, applied in order)
  void applyQT(double *c, int len) const
  {
    const int rows = qr.r;
    for (int k = 0; k < len; ++k) {
      if (tau[k] == 0.0)
        continue;
      double s = c[k];
      for (int i = k + 1; i < rows; ++i)
        s += qr(i, k) * c[i];
      s *= tau[k];
      c[k] -= s;
      for (int i = k + 1; i < rows; ++i)
        c[i] -= s * qr(i, k);
    }
  }
  // c := Q c  (first `len` reflectors, applied in reverse order)
  void applyQ(double *c, int len) const
  {
    const int rows = qr.r;
    for (int k = len - 1; k >= 0; --k) {
      if (tau[k] == 0.0)
        continue;
      double s = c[k];
      for (int i = k + 1; i < rows; ++i)
        s += qr(i, k) * c[i];
      s *= tau[k];
      c[k] -= s;
      for (int i = k + 1; i < rows; ++i)
        c[i] -= s * qr(i, k);
    }
  }

  // least-squares / square solve A x = b for each column of B
  Mat solve(const Mat &B) const
  {
    const int rows = qr.r, cols = qr.c, np = nonzeroPivots;
    Mat       X(cols, B.c);
    if (np == 0)
      return X;
    std::vector<double> c(rows);
    for (int j = 0; j < B.c; ++j) {
      std::copy(B.col(j), B.col(j) + rows, c.begin());
      applyQT(c.data(), np);
      for (int i = np - 1; i >= 0; --i) { // R(0:np,0:np) y = c
        double s = c[i];
        for (int k = i + 1; k < np; ++k)
          s -= qr(i, k) * c[k];
        c[i] = s / qr(i, i);
      }
      for (int i = 0; i < np; ++i)
        X(perm[i], j) = c[i];
      for (int i = np; i < cols; ++i)
        X(perm[i], j) = 0.0;
    }
    return X;
  }

  // minimum-norm solution of A^T x = b  (A is rows x cols, b has `cols` rows, x has `rows` rows)
  Mat solveTransposed(const Mat &B) const
  {
    const int rows = qr.r, cols = qr.c, np = nonzeroPivots;
    Mat       X(rows, B.c);
    if (np == 0)
      return X;
    std::vector<double> c(std::max(rows, cols));
    for (int j = 0; j < B.c; ++j) {
      for (int i = 0; i < cols; ++i) // c = P^T b
        c[i] = B(perm[i], j);
      for (int i = 0; i < np; ++i) { // R^T y = c (forward substitution)
        double s = c[i];
        for (int k = 0; k < i; ++k)
          s -= qr(k, i) * c[k];
        c[i] = s / qr(i, i);
      }
      double *x = X.col(j);
      for (int i = 0; i < np; ++i)
        x[i] = c[i];
      for (int i = np; i < rows; ++i)
        x[i] = 0.0;
      applyQ(x, np);
    }
    return X;
  }
};

// ---------------------------------------------------------------------------------------------
// Singular values (descending) of a small dense matrix by one-sided (Hestenes) Jacobi rotations.
// Stands in for Eigen::JacobiSVD at RadialBasisFctSolver.hpp:389-392, where only
// sigma_max / max(sigma_min, 1e-14) is consumed.
inline std::vector<double> singularValues(Mat A)
{
  // work on the matrix with fewer columns than rows
  if (A.c > A.r) {
    Mat T(A.c, A.r);
    for (int i = 0; i < A.r; ++i)
      for (int j = 0; j < A.c; ++j)
        T(j, i) = A(i, j);
    A = std::move(T);
  }
  const int m = A.r, n = A.c;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < m; ++i) {
          alpha += A(i, p) * A(i, p);
          beta += A(i, q) * A(i, q);
          gamma += A(i, p) * A(i, q);
        }
        if (gamma == 0.0 || std::abs(gamma) <= 1e-300 || std::abs(gamma) <= std::numeric_limits<double>::epsilon() * std::sqrt(alpha * beta))
          continue;
        rotated           = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t    = (zeta >= 0 ? 1.0 : -1.0) / (std::abs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double cs   = 1.0 / std::sqrt(1.0 + t * t);
        const double sn   = cs * t;
        for (int i = 0; i < m; ++i) {
          const double ap = A(i, p), aq = A(i, q);
          A(i, p) = cs * ap - sn * aq;
          A(i, q) = sn * ap + cs * aq;
        }
      }
    if (!rotated)
      break;
  }
  std::vector<double> sv(n);
  for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int i = 0; i < m; ++i)
      s += A(i, j) * A(i, j);
    sv[j] = std::sqrt(s);
  }
  std::sort(sv.begin(), sv.end(), std::greater<double>());
  return sv;
}

} // namespace orc
