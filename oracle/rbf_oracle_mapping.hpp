// rbf_oracle_mapping.hpp — CPU ORACLE (test infrastructure, NOT the product), part 2:
// spatial queries, the per-system RBF solver, clustering, partition-of-unity and global mappings.
// See rbf_oracle.hpp for the scope statement. All file:line citations are relative to
// /root/reference/src/.
#pragma once
#include <memory>

#include "rbf_oracle.hpp"

namespace orc {

// ---------------------------------------------------------------------------------------------
// Uniform-grid point index: replaces the Boost.Geometry R*-tree for exactly the four queries the
// hot path issues (query/Index.cpp:171-195 nearest / kNN, :225-253 sphere / any-in-sphere,
// :369-391 bounds). Semantics kept: sphere membership is STRICT  sqrt(d2) < radius  on all three
// stored components (Index.cpp:235, RTreeAdapter.hpp:111). Nearest-neighbour ties are
// implementation-defined in the reference; here the lowest vertex ID wins.
class PointGrid {
public:
  PointGrid() = default;
  PointGrid(const double *xyz, int n)
  {
    build(xyz, n);
  }

  void build(const double *xyz, int n)
  {
    _xyz = xyz;
    _n   = n;
    for (int d = 0; d < 3; ++d) {
      _min[d] = std::numeric_limits<double>::max();
      _max[d] = std::numeric_limits<double>::lowest();
    }
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) {
        _min[d] = std::min(_min[d], xyz[3 * i + d]);
        _max[d] = std::max(_max[d], xyz[3 * i + d]);
      }
    if (n == 0) {
      _cells = {{1, 1, 1}};
      _start.assign(2, 0);
      return;
    }
    // target ~2 points per cell over the non-degenerate axes
    double vol  = 1.0;
    int    live = 0;
    for (int d = 0; d < 3; ++d)
      if (_max[d] > _min[d]) {
        vol *= (_max[d] - _min[d]);
        ++live;
      }
    double h = live ? std::pow(vol * 2.0 / std::max(1, n), 1.0 / live) : 1.0;
    if (!(h > 0))
      h = 1.0;
    for (int d = 0; d < 3; ++d) {
      const double e = _max[d] - _min[d];
      long         c = static_cast<long>(std::floor(e / h)) + 1;
      c             = std::max(1L, std::min(c, 1024L));
      _cells[d]     = static_cast<int>(c);
      _h[d]         = e > 0 ? e / c : 1.0;
      // make sure the max coordinate falls into the last cell
    }
    const size_t nc = static_cast<size_t>(_cells[0]) * _cells[1] * _cells[2];
    _start.assign(nc + 1, 0);
    std::vector<int> cellOf(n);
    for (int i = 0; i < n; ++i) {
      cellOf[i] = cellIndex(xyz + 3 * i);
      ++_start[cellOf[i] + 1];
    }
    for (size_t c = 0; c < nc; ++c)
      _start[c + 1] += _start[c];
    _ids.resize(n);
    std::vector<int> fill(_start.begin(), _start.end() - 1);
    for (int i = 0; i < n; ++i) // ascending IDs inside each cell
      _ids[fill[cellOf[i]]++] = i;
  }

  int           size() const { return _n; }
  const double *coords(int i) const { return _xyz + 3 * i; }
  const Vec3   &minCorner() const { return _min; }
  const Vec3   &maxCorner() const { return _max; }

  // Index.cpp:225-238 — all vertices with distance < radius, returned SORTED ascending
  // (the caller puts them in a flat_set anyway, SphericalVertexCluster.hpp:161-162).
  std::vector<int> verticesInsideSphere(const double *c, double radius) const
  {
    std::vector<int> res;
    forCellsInBox(c, radius, [&](int cell) {
      for (int k = _start[cell]; k < _start[cell + 1]; ++k) {
        const int i = _ids[k];
        if (std::sqrt(squaredDifference(c, coords(i))) < radius)
          res.push_back(i);
      }
      return true;
    });
    std::sort(res.begin(), res.end());
    return res;
  }

  // Index.cpp:240-253
  bool isAnyVertexInsideSphere(const double *c, double radius) const
  {
    bool found = false;
    forCellsInBox(c, radius, [&](int cell) {
      for (int k = _start[cell]; k < _start[cell + 1]; ++k)
        if (std::sqrt(squaredDifference(c, coords(_ids[k]))) < radius) {
          found = true;
          return false;
        }
      return true;
    });
    return found;
  }

  // Index.cpp:184-195 — k nearest vertices (any order); fewer if the mesh is smaller
  std::vector<int> closestVertices(const double *c, int k) const
  {
    k = std::min(k, _n);
    std::vector<std::pair<double, int>> cand;
    if (k <= 0)
      return {};
    int ci[3];
    cellCoords(c, ci);
    const int maxRing = std::max({_cells[0], _cells[1], _cells[2]});
    for (int ring = 0; ring <= maxRing; ++ring) {
      visitRing(ci, ring, [&](int cell) {
        for (int p = _start[cell]; p < _start[cell + 1]; ++p)
          cand.emplace_back(squaredDifference(c, coords(_ids[p])), _ids[p]);
      });
      if (static_cast<int>(cand.size()) >= k) {
        // every unvisited point is farther than ring * min cell width from c (c may be outside)
        const double safe = ringSafeDistance(c, ci, ring);
        std::nth_element(cand.begin(), cand.begin() + (k - 1), cand.end());
        if (cand[k - 1].first <= safe * safe)
          break;
      }
    }
    std::sort(cand.begin(), cand.end());
    std::vector<int> res(k);
    for (int i = 0; i < k; ++i)
      res[i] = cand[i].second;
    return res;
  }

  // Index.cpp:171-182
  int closestVertex(const double *c) const
  {
    return closestVertices(c, 1)[0];
  }

private:
  const double    *_xyz = nullptr;
  int              _n   = 0;
  Vec3             _min{}, _max{}, _h{{1, 1, 1}};
  std::array<int, 3> _cells{{1, 1, 1}};
  std::vector<int> _start, _ids;

  void cellCoords(const double *p, int *ci) const
  {
    for (int d = 0; d < 3; ++d) {
      long c = static_cast<long>(std::floor((p[d] - _min[d]) / _h[d]));
      ci[d]  = static_cast<int>(std::max(0L, std::min<long>(c, _cells[d] - 1)));
    }
  }
  int cellIndex(const double *p) const
  {
    int ci[3];
    cellCoords(p, ci);
    return (ci[0] * _cells[1] + ci[1]) * _cells[2] + ci[2];
  }

  template <typename F>
  void forCellsInBox(const double *c, double radius, F &&f) const
  {
    if (_n == 0)
      return;
    int lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
      const double a = c[d] - radius, b = c[d] + radius;
      if (b < _min[d] || a > _max[d])
        return;
      long l = static_cast<long>(std::floor((a - _min[d]) / _h[d])) - 1;
      long h = static_cast<long>(std::floor((b - _min[d]) / _h[d])) + 1;
      lo[d]  = static_cast<int>(std::max(0L, std::min<long>(l, _cells[d] - 1)));
      hi[d]  = static_cast<int>(std::max(0L, std::min<long>(h, _cells[d] - 1)));
    }
    for (int x = lo[0]; x <= hi[0]; ++x)
      for (int y = lo[1]; y <= hi[1]; ++y)
        for (int z = lo[2]; z <= hi[2]; ++z)
          if (!f((x * _cells[1] + y) * _cells[2] + z))
            return;
  }

  // visits all cells with Chebyshev cell-distance == ring from ci (clipped to the grid)
  template <typename F>
  void visitRing(const int *ci, int ring, F &&f) const
  {
    const int x0 = ci[0] - ring, x1 = ci[0] + ring, y0 = ci[1] - ring, y1 = ci[1] + ring, z0 = ci[2] - ring, z1 = ci[2] + ring;
    for (int x = std::max(0, x0); x <= std::min(_cells[0] - 1, x1); ++x)
      for (int y = std::max(0, y0); y <= std::min(_cells[1] - 1, y1); ++y) {
        const bool edgeXY = (x == x0 || x == x1 || y == y0 || y == y1);
        if (edgeXY) {
          for (int z = std::max(0, z0); z <= std::min(_cells[2] - 1, z1); ++z)
            f((x * _cells[1] + y) * _cells[2] + z);
        } else {
          if (z0 >= 0)
            f((x * _cells[1] + y) * _cells[2] + z0);
          if (z1 <= _cells[2] - 1 && z1 != z0)
            f((x * _cells[1] + y) * _cells[2] + z1);
        }
      }
  }

  // lower bound for the distance from c to any point in a cell outside the visited block
  double ringSafeDistance(const double *c, const int *ci, int ring) const
  {
    double safe = std::numeric_limits<double>::max();
    bool   any  = false;
    for (int d = 0; d < 3; ++d) {
      // distance from c to the faces of the visited block along d, if cells exist beyond them
      const int lo = ci[d] - ring, hi = ci[d] + ring;
      if (lo > 0) {
        const double face = _min[d] + lo * _h[d];
        safe              = std::min(safe, std::max(0.0, c[d] - face));
        any               = true;
      }
      if (hi < _cells[d] - 1) {
        const double face = _min[d] + (hi + 1) * _h[d];
        safe              = std::min(safe, std::max(0.0, face - c[d]));
        any               = true;
      }
    }
    return any ? safe : std::numeric_limits<double>::max();
  }
};

// ---------------------------------------------------------------------------------------------
// RadialBasisFctSolver.hpp:140-162 — columns [1, x_active...] starting at `startCol`
inline void fillPolynomialEntries(Mat &M, const double *xyz, const std::vector<int> &ids, int startCol, const std::array<bool, 3> &active)
{
  for (int i = 0; i < static_cast<int>(ids.size()); ++i) {
    M(i, startCol)  = 1.0;
    const double *u = xyz + 3 * static_cast<size_t>(ids[i]);
    int           k = 0;
    for (int d = 0; d < 3; ++d)
      if (active[d]) {
        M(i, startCol + 1 + k) = u[d];
        ++k;
      }
  }
}

// RadialBasisFctSolver.hpp:115-137 — deactivate the active axis with the smallest extent
inline void reduceActiveAxis(const double *xyz, const std::vector<int> &ids, std::array<bool, 3> &axis)
{
  std::array<std::pair<int, double>, 3> diff;
  for (int d = 0; d < 3; ++d) {
    if (!axis[d]) {
      diff[d] = {d, std::numeric_limits<double>::max()};
    } else {
      double lo = std::numeric_limits<double>::max(), hi = std::numeric_limits<double>::lowest();
      for (int id : ids) {
        lo = std::min(lo, xyz[3 * static_cast<size_t>(id) + d]);
        hi = std::max(hi, xyz[3 * static_cast<size_t>(id) + d]);
      }
      diff[d] = {d, std::abs(hi - lo)};
    }
  }
  std::sort(diff.begin(), diff.end(), [](const auto &a, const auto &b) { return a.second < b.second; });
  axis[diff[0].first] = false;
}

// ---------------------------------------------------------------------------------------------
// One RBF system: restates mapping::RadialBasisFctSolver (RadialBasisFctSolver.hpp:340-468).
class RbfSolver {
public:
  RbfSolver() = default;

  // ctor: :340-411. inXYZ/outXYZ are the full meshes, inIDs/outIDs select the vertices.
  RbfSolver(const Rbf &f, const double *inXYZ, const std::vector<int> &inIDs, const double *outXYZ, const std::vector<int> &outIDs,
            const std::array<bool, 3> &deadAxis, Polynomial polynomial, const std::string &inName = "in", const std::string &outName = "out")
  {
    std::array<bool, 3> active{{!deadAxis[0], !deadAxis[1], !deadAxis[2]}};
    const int           dead       = static_cast<int>(std::count(active.begin(), active.end(), false));
    const int           polyparams = polynomial == Polynomial::ON ? 1 + 3 - dead : 0;
    const int           n = static_cast<int>(inIDs.size()), m = static_cast<int>(outIDs.size());
    _spd = f.isStrictlyPositiveDefinite();

    // buildMatrixCLU :164-207 (upper triangle evaluated, mirrored)
    Mat C(n + polyparams, n + polyparams);
    for (int i = 0; i < n; ++i) {
      const double *u = inXYZ + 3 * static_cast<size_t>(inIDs[i]);
      for (int j = i; j < n; ++j) {
        const double *v = inXYZ + 3 * static_cast<size_t>(inIDs[j]);
        C(i, j)         = f(std::sqrt(squaredDifference(u, v, active)));
      }
    }
    if (polynomial == Polynomial::ON)
      fillPolynomialEntries(C, inXYZ, inIDs, n, active);
    for (int j = 0; j < C.c; ++j)
      for (int i = j + 1; i < C.r; ++i)
        C(i, j) = C(j, i);

    bool ok;
    if (_spd) {
      _llt.compute(std::move(C));
      ok = _llt.ok;
    } else {
      _qrC.compute(std::move(C));
      ok = _qrC.isInvertible();
    }
    if (!ok)
      throw Error("The interpolation matrix of the RBF mapping from mesh \"" + inName + "\" to mesh \"" + outName +
                  "\" is not invertable. This means that the mapping problem is not well-posed. "
                  "Please check if your coupling meshes are correct (e.g. no vertices are duplicated) or reconfigure "
                  "your basis-function (e.g. reduce the support-radius).");

    // buildMatrixA :209-242
    _A = Mat(m, n + polyparams);
    for (int i = 0; i < m; ++i) {
      const double *u = outXYZ + 3 * static_cast<size_t>(outIDs[i]);
      for (int j = 0; j < n; ++j) {
        const double *v = inXYZ + 3 * static_cast<size_t>(inIDs[j]);
        _A(i, j)        = f(std::sqrt(squaredDifference(u, v, active)));
      }
    }
    if (polynomial == Polynomial::ON)
      fillPolynomialEntries(_A, outXYZ, outIDs, n, active);

    // separated polynomial :377-410
    _localActive = active;
    if (polynomial == Polynomial::SEPARATE) {
      int polyParams = numberOfPolynomials();
      while (true) {
        _Q = Mat(n, polyParams);
        fillPolynomialEntries(_Q, inXYZ, inIDs, 0, _localActive);
        const auto   sv   = singularValues(_Q);
        const double cond = sv.front() / std::max(sv.back(), NUMERICAL_ZERO_DIFFERENCE);
        if (cond > 1e5) {
          reduceActiveAxis(inXYZ, inIDs, _localActive);
          polyParams = numberOfPolynomials();
        } else {
          break;
        }
      }
      _V = Mat(m, polyParams);
      fillPolynomialEntries(_V, outXYZ, outIDs, 0, _localActive);
      _qrQ.compute(_Q);
    }
    _n = n + polyparams;
    _m = m;
  }

  int inputSize() const { return _n; }
  int outputSize() const { return _m; }
  int numberOfPolynomials() const { return 1 + static_cast<int>(std::count(_localActive.begin(), _localActive.end(), true)); }
  const std::array<bool, 3> &localActiveAxis() const { return _localActive; }

  Mat solveC(const Mat &B) const { return _spd ? _llt.solve(B) : _qrC.solve(B); }

  // :441-468  (in: inputSize x dims, modified in place like the reference)
  Mat solveConsistent(Mat &in, Polynomial polynomial) const
  {
    Mat beta;
    if (polynomial == Polynomial::SEPARATE) {
      beta        = _qrQ.solve(in);
      const Mat P = matmul(_Q, beta);
      for (size_t i = 0; i < in.a.size(); ++i)
        in.a[i] -= P.a[i];
    }
    const Mat p   = solveC(in);
    Mat       out = matmul(_A, p);
    if (polynomial == Polynomial::SEPARATE) {
      const Mat P = matmul(_V, beta);
      for (size_t i = 0; i < out.a.size(); ++i)
        out.a[i] += P.a[i];
    }
    return out;
  }

  // :413-438  (in: outputSize x dims)
  Mat solveConservative(const Mat &in, Polynomial polynomial) const
  {
    const Mat Au  = matmulT(_A, in);
    Mat       out = solveC(Au);
    if (polynomial == Polynomial::SEPARATE) {
      Mat       eps = matmulT(_V, in);
      const Mat Qm  = matmulT(_Q, out);
      for (size_t i = 0; i < eps.a.size(); ++i)
        eps.a[i] = -(eps.a[i] - Qm.a[i]); // -epsilon
      const Mat sigma = _qrQ.solveTransposed(eps);
      for (size_t i = 0; i < out.a.size(); ++i)
        out.a[i] -= sigma.a[i];
    }
    return out;
  }

  // ---- just-in-time mapping (RadialBasisFctSolver.hpp:470-573) ----
  // computeCacheData :470-482
  void computeCacheData(Mat &in, Polynomial polynomial, Mat &polyOut, Mat &coeffsOut) const
  {
    if (polynomial == Polynomial::SEPARATE) {
      polyOut     = _qrQ.solve(in);
      const Mat P = matmul(_Q, polyOut);
      for (size_t i = 0; i < in.a.size(); ++i)
        in.a[i] -= P.a[i];
    }
    coeffsOut = solveC(in);
  }

  // interpolateAt :484-520 (all three axes in the distance, "we ignore the dead axis here")
  void interpolateAt(const double *v, const Mat &poly, const Mat &coeffs, const Rbf &f, const double *inXYZ, const std::vector<int> &inIDs, double *result) const
  {
    const int dims = coeffs.c;
    for (int d = 0; d < dims; ++d)
      result[d] = 0.0;
    for (size_t j = 0; j < inIDs.size(); ++j) {
      const double eval = f(std::sqrt(squaredDifference(v, inXYZ + 3 * static_cast<size_t>(inIDs[j]))));
      for (int d = 0; d < dims; ++d)
        result[d] += eval * coeffs(static_cast<int>(j), d);
    }
    if (!poly.a.empty()) {
      for (int d = 0; d < dims; ++d)
        result[d] += 1 * poly(0, d);
      int k = 1;
      for (int ax = 0; ax < 3; ++ax)
        if (_localActive[ax]) {
          for (int d = 0; d < dims; ++d)
            result[d] += v[ax] * poly(k, d);
          ++k;
        }
    }
  }

  // addWriteDataToCache :522-560
  void addWriteDataToCache(const double *v, const double *load, Mat &epsilon, Mat &Au, const Rbf &f, const double *inXYZ, const std::vector<int> &inIDs) const
  {
    const int dims = Au.c;
    for (size_t j = 0; j < inIDs.size(); ++j) {
      const double eval = f(std::sqrt(squaredDifference(v, inXYZ + 3 * static_cast<size_t>(inIDs[j]))));
      for (int d = 0; d < dims; ++d)
        Au(static_cast<int>(j), d) += eval * load[d];
    }
    if (!epsilon.a.empty()) {
      for (int d = 0; d < dims; ++d)
        epsilon(0, d) += 1 * load[d];
      int k = 1;
      for (int ax = 0; ax < 3; ++ax)
        if (_localActive[ax]) {
          for (int d = 0; d < dims; ++d)
            epsilon(k, d) += v[ax] * load[d];
          ++k;
        }
    }
  }

  // evaluateConservativeCache :562-573
  Mat evaluateConservativeCache(Mat &epsilon, const Mat &Au) const
  {
    Mat out = solveC(Au);
    if (!epsilon.a.empty()) {
      const Mat Qm = matmulT(_Q, out);
      for (size_t i = 0; i < epsilon.a.size(); ++i)
        epsilon.a[i] -= Qm.a[i];
      Mat neg = epsilon;
      for (auto &e : neg.a)
        e = -e;
      const Mat sigma = _qrQ.solveTransposed(neg);
      for (size_t i = 0; i < out.a.size(); ++i)
        out.a[i] -= sigma.a[i];
    }
    return out;
  }

  const Mat &matrixA() const { return _A; }
  const Mat &choleskyL() const { return _llt.L; }

private:
  bool                _spd = true;
  Cholesky            _llt;
  ColPivQR            _qrC;
  ColPivQR            _qrQ;
  Mat                 _A, _Q, _V;
  std::array<bool, 3> _localActive{{true, true, true}};
  int                 _n = 0, _m = 0;
};

// ---------------------------------------------------------------------------------------------
// Clustering: restates impl::createClustering and helpers (impl/CreateClustering.hpp:36-505).
struct Clustering {
  double              radius = 0.0;
  std::vector<double> centers; // 3 doubles per centre
  int                 count() const { return static_cast<int>(centers.size() / 3); }
};

// impl/CreateClustering.hpp:223-278
inline double estimateClusterRadius(unsigned verticesPerCluster, const PointGrid &in, int dim, const Vec3 &bbMin, const Vec3 &bbMax)
{
  std::vector<int> samples;
  Vec3             c{};
  for (int d = 0; d < 3; ++d)
    c[d] = (bbMax[d] + bbMin[d]) * 0.5;
  if (dim == 2)
    c[2] = 0.0;
  samples.push_back(in.closestVertex(c.data()));
  for (int d = 0; d < dim; ++d) {
    Vec3         s    = c;
    const double edge = std::abs(bbMax[d] - bbMin[d]);
    s[d] -= edge * 0.25;
    samples.push_back(in.closestVertex(s.data()));
    s[d] += edge * 0.5;
    samples.push_back(in.closestVertex(s.data()));
  }
  std::vector<double> radii;
  for (int s : samples) {
    const auto knn   = in.closestVertices(in.coords(s), static_cast<int>(verticesPerCluster));
    double     maxSq = 0.0;
    for (int i : knn)
      maxSq = std::max(maxSq, squaredDifference(in.coords(i), in.coords(s)));
    radii.push_back(std::sqrt(maxSq));
  }
  const size_t middle = radii.size() / 2;
  std::nth_element(radii.begin(), radii.begin() + middle, radii.end());
  return radii[middle];
}

// impl/CreateClustering.hpp:302-505. `outIsJustInTime` mirrors outMesh->isJustInTime().
inline Clustering createClustering(const PointGrid &in, const PointGrid &out, int dim, double relativeOverlap, unsigned verticesPerCluster,
                                   bool projectClustersToInput, bool outIsJustInTime = false)
{
  Clustering res;
  if (in.size() == 0 || (out.size() == 0 && !outIsJustInTime))
    return res; // :314-316
  const bool startGridAtEdge = projectClustersToInput;
  Vec3       bbMin = in.minCorner(), bbMax = in.maxCorner(); // :334
  auto       edge = [&](int d) { return std::abs(bbMax[d] - bbMin[d]); };

  // single cluster fallback :352-359
  if (static_cast<unsigned>(in.size()) < verticesPerCluster * 2) {
    double cRadius = 0.0;
    for (int d = 0; d < dim; ++d)
      cRadius = d == 0 ? (bbMax[d] - bbMin[d]) : std::max(cRadius, bbMax[d] - bbMin[d]);
    if (cRadius == 0)
      cRadius = 1;
    res.radius = cRadius * 2;
    res.centers.assign(3, 0.0);
    for (int d = 0; d < dim; ++d)
      res.centers[d] = (bbMax[d] + bbMin[d]) * 0.5;
    return res;
  }

  const double clusterRadius         = estimateClusterRadius(verticesPerCluster, in, dim, bbMin, bbMax);
  const double maximumCenterDistance = std::sqrt(4. / dim) * clusterRadius * (1 - relativeOverlap);

  std::array<unsigned, 3> nGlobal{{1, 1, 1}};
  for (int d = 0; d < dim; ++d)
    nGlobal[d] = static_cast<unsigned>(std::ceil(std::max(1., edge(d) / maximumCenterDistance)));

  std::vector<double> distances(dim), start(dim);
  for (int d = 0; d < dim; ++d) {
    distances[d] = edge(d) / nGlobal[d];
    start[d]     = bbMin[d];
  }
  if (startGridAtEdge) {
    for (int d = 0; d < dim; ++d)
      if (edge(d) > NUMERICAL_ZERO_DIFFERENCE)
        nGlobal[d] += 1;
  } else {
    for (int d = 0; d < dim; ++d)
      start[d] = start[d] + 0.5 * distances[d];
  }
  std::array<unsigned, 3> nLocal{{1, 1, 1}};
  for (int d = 0; d < dim; ++d) {
    if (distances[d] > 0) {
      start[d] += std::ceil(((bbMin[d] - start[d]) / distances[d]) - NUMERICAL_ZERO_DIFFERENCE) * distances[d];
      nLocal[d] = static_cast<unsigned>(1 + std::floor(((bbMax[d] - start[d]) / distances[d]) + NUMERICAL_ZERO_DIFFERENCE));
    }
  }
  const size_t        total = static_cast<size_t>(nLocal[0]) * nLocal[1] * nLocal[2];
  std::vector<double> centers(total * 3, 0.0);
  std::vector<char>   tagged(total, 0);
  auto                lin = [&](unsigned x, unsigned y, unsigned z) { return (static_cast<size_t>(x) * nLocal[1] + y) * nLocal[2] + z; };

  // Step 7 :438-463 — coordinates grow by repeated addition, exactly like the reference loops
  {
    std::vector<double> cc(3, 0.0);
    cc[0] = start[0];
    for (unsigned x = 0; x < nLocal[0]; ++x, cc[0] += distances[0]) {
      cc[1] = start[1];
      for (unsigned y = 0; y < nLocal[1]; ++y, cc[1] += distances[1]) {
        if (dim == 2) {
          const size_t id     = lin(x, y, 0);
          centers[3 * id]     = cc[0];
          centers[3 * id + 1] = cc[1];
          centers[3 * id + 2] = 0.0;
        } else {
          cc[2] = start[2];
          for (unsigned z = 0; z < nLocal[2]; ++z, cc[2] += distances[2]) {
            const size_t id     = lin(x, y, z);
            centers[3 * id]     = cc[0];
            centers[3 * id + 1] = cc[1];
            centers[3 * id + 2] = cc[2];
          }
        }
      }
    }
  }

  auto tagEmpty = [&](const PointGrid &mesh) { // :132-140
    for (size_t i = 0; i < total; ++i)
      if (!tagged[i] && !mesh.isAnyVertexInsideSphere(&centers[3 * i], clusterRadius))
        tagged[i] = 1;
  };
  tagEmpty(in);
  if (!outIsJustInTime)
    tagEmpty(out);

  if (projectClustersToInput) {
    // :150-160
    for (size_t i = 0; i < total; ++i)
      if (!tagged[i]) {
        const int v = in.closestVertex(&centers[3 * i]);
        for (int d = 0; d < 3; ++d)
          centers[3 * i + d] = in.coords(v)[d];
      }
    const double threshold = 0.4 * (*std::min_element(distances.begin(), distances.end()));
    // :173-195 with :54-82 (index-neighbour offsets) and :96-122 (greedy, order dependent)
    std::vector<long> offsets;
    const bool        twoD = (nLocal[2] == 1);
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = (twoD ? 0 : -1); dz <= (twoD ? 0 : 1); ++dz) {
          if (dx == 0 && dy == 0 && dz == 0)
            continue;
          offsets.push_back((static_cast<long>(dx) * nLocal[1] + dy) * static_cast<long>(nLocal[2]) + dz);
        }
    const double thr2 = pow_int<2>(threshold);
    for (size_t i = 0; i < total; ++i) {
      if (tagged[i])
        continue;
      bool hit = false;
      for (long off : offsets) {
        if (off == 0)
          continue;
        const long nb = static_cast<long>(i) + off;
        if (nb >= 0 && nb < static_cast<long>(total) && !tagged[nb] && squaredDifference(&centers[3 * i], &centers[3 * nb]) < thr2) {
          hit = true;
          break;
        }
      }
      if (hit)
        tagged[i] = 1;
    }
    if (!outIsJustInTime)
      tagEmpty(out);
  }

  for (size_t i = 0; i < total; ++i)
    if (!tagged[i])
      res.centers.insert(res.centers.end(), centers.begin() + 3 * i, centers.begin() + 3 * i + 3);
  if (res.centers.empty())
    throw Error("Too many vertices have been filtered out.");
  res.radius = clusterRadius;
  return res;
}

// ---------------------------------------------------------------------------------------------
// helper: run f(i) for i in [0,n) on `threads` host threads (static contiguous split). Used only
// for the CPU-N baseline that emulates one MPI rank per core.
template <typename F>
inline void parallelFor(size_t n, int threads, F &&f)
{
  threads = std::max(1, std::min<int>(threads, static_cast<int>(std::max<size_t>(1, n))));
  if (threads == 1) {
    for (size_t i = 0; i < n; ++i)
      f(i, 0);
    return;
  }
  std::vector<std::thread> pool;
  std::vector<std::string> errors(threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t]() {
      const size_t b = n * t / threads, e = n * (t + 1) / threads;
      try {
        for (size_t i = b; i < e; ++i)
          f(i, t);
      } catch (const std::exception &ex) {
        errors[t] = ex.what();
      }
    });
  for (auto &th : pool)
    th.join();
  for (auto &e : errors)
    if (!e.empty())
      throw Error(e);
}

inline void atomicAddDouble(double *addr, double v)
{
  auto    *p   = reinterpret_cast<uint64_t *>(addr);
  uint64_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (true) {
    double cur;
    std::memcpy(&cur, &old, 8);
    cur += v;
    uint64_t nw;
    std::memcpy(&nw, &cur, 8);
    if (__atomic_compare_exchange_n(p, &old, nw, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
      return;
  }
}

// ---------------------------------------------------------------------------------------------
// Partition-of-unity mapping: restates PartitionOfUnityMapping (PartitionOfUnityMapping.hpp:181-380)
// with SphericalVertexCluster (impl/SphericalVertexCluster.hpp:139-240, 296-344).
class PumMapping {
public:
  struct Cluster {
    Vec3                center;
    std::vector<int>    inIDs, outIDs;
    RbfSolver           solver;
    std::vector<double> weights;
    bool                built = true; // false: placeholder of a cluster outside the sampled region (computeMappingForSamples)
  };

  PumMapping(Constraint constraint, int dim, const Rbf &f, Polynomial polynomial, unsigned verticesPerCluster, double relativeOverlap, bool projectToInput)
      : _constraint(constraint), _dim(dim), _f(f), _poly(polynomial), _vpc(verticesPerCluster), _overlap(relativeOverlap), _project(projectToInput)
  {
    if (polynomial == Polynomial::ON)
      throw Error("Integrated polynomial is not supported for partition of unity data mappings.");
  }

  // computeMapping :181-223 + _computeCPU :225-288. `input`/`output` are the meshes in the
  // Mapping::input()/output() sense; the swap for conservative constraints happens here (:190-198).
  // The other mesh is a just-in-time mesh (mesh::MeshConfiguration::getJustInTimeMappingMesh): consistent -> the data
  // is read at arbitrary points later (mapConsistentAt), conservative -> loads are written at arbitrary points
  // (mapConservativeAt + completeJustInTimeMapping). `meshXYZ` is the real mesh (= solver-side inMesh in both cases).
  void computeMappingJustInTime(const double *meshXYZ, int n, int threads = 1)
  {
    _jit = true;
    if (_constraint == Constraint::CONSERVATIVE)
      computeMapping(nullptr, 0, meshXYZ, n, threads);
    else
      computeMapping(meshXYZ, n, nullptr, 0, threads);
  }

  struct Cache { // impl/MappingDataCache.hpp:23-71
    int              dims = 0;
    std::vector<Mat> poly, p;
  };

  // initializeMappingDataCache :424-434 (+ MappingDataCache::resetData)
  void initializeCache(Cache &cache, int dims) const
  {
    cache.dims = dims;
    cache.poly.assign(_clusters.size(), Mat());
    cache.p.assign(_clusters.size(), Mat());
    for (size_t c = 0; c < _clusters.size(); ++c) {
      if (_poly == Polynomial::SEPARATE)
        cache.poly[c] = Mat(_clusters[c].solver.numberOfPolynomials(), dims);
      cache.p[c] = Mat(static_cast<int>(_clusters[c].inIDs.size()), dims);
    }
  }

  // updateMappingDataCache :436-448 -> SphericalVertexCluster::computeCacheData :255-270
  void updateCache(Cache &cache, const double *in) const
  {
    const int dims = cache.dims;
    for (size_t c = 0; c < _clusters.size(); ++c) {
      const Cluster &k = _clusters[c];
      Mat            f(k.solver.inputSize(), dims);
      for (size_t i = 0; i < k.inIDs.size(); ++i)
        for (int d = 0; d < dims; ++d)
          f(static_cast<int>(i), d) = in[static_cast<size_t>(k.inIDs[i]) * dims + d];
      k.solver.computeCacheData(f, _poly, cache.poly[c], cache.p[c]);
    }
  }

  // computeNormalizedWeight :290-341 for an arbitrary point
  void normalizedWeights(const double *x, std::vector<int> &ids, std::vector<double> &w) const
  {
    ids = _centerGrid->verticesInsideSphere(x, _radius);
    if (ids.empty())
      throw Error("Output vertex could not be assigned to any cluster in the rbf-pum mapping. This probably means that the meshes do not match well geometry-wise: Visualize the exported preCICE meshes to confirm. "
                  "If the meshes are fine geometry-wise, you can try to increase the number of \"vertices-per-cluster\" (default is 50), the \"relative-overlap\" (default is 0.15), or disable the option \"project-to-input\". "
                  "These options are only valid for the <mapping:rbf-pum-direct/> tag.");
    const Rbf weighting = Rbf::make(BasisFunction::WendlandC2, _radius);
    w.assign(ids.size(), 0.0);
    double sum = 0.0;
    for (size_t i = 0; i < ids.size(); ++i) {
      w[i] = weighting(std::sqrt(squaredDifference(_clusters[ids[i]].center.data(), x)));
      sum += w[i];
    }
    if (sum <= 0) {
      for (auto &e : w)
        e = 1. / w.size();
      sum = 1;
    }
    for (auto &e : w)
      e /= sum;
  }

  // mapConsistentAt :450-476; coords: nPts x 3, out: nPts x dims (vertex-major)
  void mapConsistentAt(const double *coords, int nPts, const Cache &cache, double *out) const
  {
    const int           dims = cache.dims;
    std::vector<double> res(dims);
    std::vector<int>    ids;
    std::vector<double> w;
    for (int v = 0; v < nPts; ++v) {
      for (int d = 0; d < dims; ++d)
        out[static_cast<size_t>(v) * dims + d] = 0.0;
      normalizedWeights(coords + 3 * v, ids, w);
      for (size_t i = 0; i < ids.size(); ++i) {
        const Cluster &k = _clusters[ids[i]];
        k.solver.interpolateAt(coords + 3 * v, cache.poly[ids[i]], cache.p[ids[i]], _f, _inXYZ.data(), k.inIDs, res.data());
        for (int d = 0; d < dims; ++d)
          out[static_cast<size_t>(v) * dims + d] += w[i] * res[d];
      }
    }
  }

  // mapConservativeAt :382-409; source: nPts x dims
  void mapConservativeAt(const double *coords, int nPts, const double *source, Cache &cache) const
  {
    const int           dims = cache.dims;
    std::vector<double> load(dims);
    std::vector<int>    ids;
    std::vector<double> w;
    for (int v = 0; v < nPts; ++v) {
      normalizedWeights(coords + 3 * v, ids, w);
      for (size_t i = 0; i < ids.size(); ++i) {
        const Cluster &k = _clusters[ids[i]];
        for (int d = 0; d < dims; ++d)
          load[d] = w[i] * source[static_cast<size_t>(v) * dims + d];
        k.solver.addWriteDataToCache(coords + 3 * v, load.data(), cache.poly[ids[i]], cache.p[ids[i]], _f, _inXYZ.data(), k.inIDs);
      }
    }
  }

  // completeJustInTimeMapping :411-422 -> SphericalVertexCluster::evaluateConservativeCache :242-253; out accumulates
  void completeJustInTimeMapping(Cache &cache, double *out) const
  {
    const int dims = cache.dims;
    for (size_t c = 0; c < _clusters.size(); ++c) {
      double sq = 0.0;
      for (double e : cache.p[c].a)
        sq += e * e;
      if (!(sq > 0))
        continue;
      const Cluster &k = _clusters[c];
      const Mat      r = k.solver.evaluateConservativeCache(cache.poly[c], cache.p[c]);
      for (size_t i = 0; i < k.inIDs.size(); ++i)
        for (int d = 0; d < dims; ++d)
          out[static_cast<size_t>(k.inIDs[i]) * dims + d] += r(static_cast<int>(i), d);
    }
  }

  // Test-infrastructure variant for full-size fixtures (tests/golden/make_golden_10m.py): the same mapping, but only the
  // clusters that contain one of `nSamples` vertices of the RESULT mesh (Mapping::output()) are built. The clustering and
  // the normalised weights are those of the complete mapping, so map() returns the complete mapping's values at the
  // sampled vertices (and partial sums elsewhere).
  void computeMappingForSamples(const double *inputXYZ, int nInput, const double *outputXYZ, int nOutput, const int *samples, int nSamples, int threads = 1)
  {
    _samples.assign(samples, samples + nSamples);
    _sampled = true;
    try {
      computeMapping(inputXYZ, nInput, outputXYZ, nOutput, threads);
    } catch (...) {
      _sampled = false;
      throw;
    }
    _sampled = false;
  }

  void computeMapping(const double *inputXYZ, int nInput, const double *outputXYZ, int nOutput, int threads = 1)
  {
    if (_computed)
      throw Error("Please clear the mapping before recomputing.");
    const bool cons = _constraint == Constraint::CONSERVATIVE;
    _inXYZ.assign(cons ? outputXYZ : inputXYZ, (cons ? outputXYZ : inputXYZ) + 3 * static_cast<size_t>(cons ? nOutput : nInput));
    _outXYZ.assign(cons ? inputXYZ : outputXYZ, (cons ? inputXYZ : outputXYZ) + 3 * static_cast<size_t>(cons ? nInput : nOutput));
    _nIn  = cons ? nOutput : nInput;
    _nOut = cons ? nInput : nOutput;
    PointGrid inGrid(_inXYZ.data(), _nIn), outGrid(_outXYZ.data(), _nOut);

    const Clustering cl = createClustering(inGrid, outGrid, _dim, _overlap, _vpc, _project, _jit);
    _radius             = cl.radius;
    const int nc        = cl.count();

    // SphericalVertexCluster ctor :139-182, one per centre candidate
    std::vector<Cluster> tmp(nc);
    // computeMappingForSamples: the centres whose cluster holds a sampled vertex of the result mesh (the solver-side
    // outMesh for consistent, inMesh for conservative constraints; membership radii of SphericalVertexCluster.hpp:155-157)
    std::vector<char> wanted(nc, _sampled ? 0 : 1);
    if (_sampled) {
      PointGrid     allCentres(cl.centers.data(), nc);
      const double *mesh = cons ? _inXYZ.data() : _outXYZ.data();
      const double  r    = cons ? _radius : _radius - NUMERICAL_ZERO_DIFFERENCE;
      for (int sv : _samples)
        for (int c : allCentres.verticesInsideSphere(mesh + 3 * static_cast<size_t>(sv), r))
          wanted[c] = 1;
    }
    const std::array<bool, 3> dead{{false, false, _dim == 2 ? true : false}};
    // NOTE: deadAxis is a std::vector<bool>(dim,false) in the reference (:178): for 2-D meshes only two
    // entries are transformed, the third activeAxis entry stays false (RadialBasisFctSolver.hpp:346-347).
    parallelFor(nc, threads, [&](size_t c, int) {
      Cluster &k = tmp[c];
      for (int d = 0; d < 3; ++d)
        k.center[d] = cl.centers[3 * c + d];
      if (!wanted[c]) { // placeholder: keeps its slot in the centre list (weights are normalised over ALL clusters)
        k.built = false;
        if (inGrid.isAnyVertexInsideSphere(k.center.data(), _radius))
          k.inIDs.assign(1, -1); // "not empty"
        return;
      }
      k.outIDs = outGrid.verticesInsideSphere(k.center.data(), _radius - NUMERICAL_ZERO_DIFFERENCE);
      k.inIDs  = inGrid.verticesInsideSphere(k.center.data(), _radius);
      if (k.inIDs.empty())
        return;
      k.solver = RbfSolver(_f, _inXYZ.data(), k.inIDs, _outXYZ.data(), k.outIDs, dead, _poly);
      k.weights.assign(k.outIDs.size(), 0.0);
    });
    _clusters.clear();
    for (auto &k : tmp)
      if (!k.inIDs.empty())
        _clusters.emplace_back(std::move(k));

    // weights :263-276 with computeNormalizedWeight :290-341
    std::vector<double> centerXYZ(3 * _clusters.size());
    for (size_t c = 0; c < _clusters.size(); ++c)
      for (int d = 0; d < 3; ++d)
        centerXYZ[3 * c + d] = _clusters[c].center[d];
    _centerXYZ = centerXYZ;
    _centerGrid.reset(new PointGrid(_centerXYZ.data(), static_cast<int>(_clusters.size()))); // kept for just-in-time queries (:285-287)
    PointGrid &centerGrid = *_centerGrid;
    const Rbf weighting = Rbf::make(BasisFunction::WendlandC2, _radius); // SphericalVertexCluster.hpp:146
    std::vector<char> needsWeight;
    if (_sampled) { // only the out-vertices of built clusters
      needsWeight.assign(_nOut, 0);
      for (const auto &k : _clusters)
        if (k.built)
          for (int v : k.outIDs)
            needsWeight[v] = 1;
    }
    parallelFor(_nOut, threads, [&](size_t v, int) {
      if (_sampled && !needsWeight[v])
        return;
      const double    *x   = &_outXYZ[3 * v];
      std::vector<int> ids = centerGrid.verticesInsideSphere(x, _radius);
      if (ids.empty())
        throw Error("Output vertex " + std::to_string(v) + " could not be assigned to any cluster in the rbf-pum mapping. This probably means that the meshes do not match well geometry-wise: Visualize the exported preCICE meshes to confirm. "
                                                            "If the meshes are fine geometry-wise, you can try to increase the number of \"vertices-per-cluster\" (default is 50), the \"relative-overlap\" (default is 0.15), or disable the option \"project-to-input\". "
                                                            "These options are only valid for the <mapping:rbf-pum-direct/> tag.");
      std::vector<double> w(ids.size());
      double              sum = 0.0;
      for (size_t i = 0; i < ids.size(); ++i) {
        w[i] = weighting(std::sqrt(squaredDifference(_clusters[ids[i]].center.data(), x))); // computeWeight :337-344
        sum += w[i];
      }
      if (sum <= 0) {
        for (auto &e : w)
          e = 1. / w.size();
        sum = 1;
      }
      for (size_t i = 0; i < ids.size(); ++i) {
        Cluster &k  = _clusters[ids[i]];
        if (!k.built)
          continue;
        auto     it = std::lower_bound(k.outIDs.begin(), k.outIDs.end(), static_cast<int>(v)); // setNormalizedWeight :184-199
        if (it != k.outIDs.end() && *it == static_cast<int>(v))
          k.weights[it - k.outIDs.begin()] = w[i] / sum;
      }
    });
    _computed = true;
  }

  // mapConsistent :362-380 -> SphericalVertexCluster::mapConsistent :296-335
  void mapConsistent(const double *in, int dims, double *out, int threads = 1) const
  {
    parallelFor(_clusters.size(), threads, [&](size_t c, int) {
      const Cluster &k = _clusters[c];
      if (!k.built)
        return;
      const int      n = static_cast<int>(k.inIDs.size());
      Mat            f(k.solver.inputSize(), dims);
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < dims; ++d)
          f(i, d) = in[static_cast<size_t>(k.inIDs[i]) * dims + d];
      const Mat r = k.solver.solveConsistent(f, _poly);
      for (size_t i = 0; i < k.outIDs.size(); ++i)
        for (int d = 0; d < dims; ++d) {
          double *dst = &out[static_cast<size_t>(k.outIDs[i]) * dims + d];
          if (threads > 1)
            atomicAddDouble(dst, r(static_cast<int>(i), d) * k.weights[i]);
          else
            *dst += r(static_cast<int>(i), d) * k.weights[i];
        }
    });
  }

  // mapConservative :343-360 -> SphericalVertexCluster::mapConservative :201-240
  void mapConservative(const double *in, int dims, double *out, int threads = 1) const
  {
    parallelFor(_clusters.size(), threads, [&](size_t c, int) {
      const Cluster &k = _clusters[c];
      if (!k.built)
        return;
      Mat            u(k.solver.outputSize(), dims);
      for (size_t i = 0; i < k.outIDs.size(); ++i)
        for (int d = 0; d < dims; ++d)
          u(static_cast<int>(i), d) = in[static_cast<size_t>(k.outIDs[i]) * dims + d] * k.weights[i];
      const Mat r = k.solver.solveConservative(u, _poly);
      for (size_t i = 0; i < k.inIDs.size(); ++i)
        for (int d = 0; d < dims; ++d) {
          double *dst = &out[static_cast<size_t>(k.inIDs[i]) * dims + d];
          if (threads > 1)
            atomicAddDouble(dst, r(static_cast<int>(i), d));
          else
            *dst += r(static_cast<int>(i), d);
        }
    });
  }

  // Mapping::map dispatch, Mapping.cpp:158-173 (scaled-consistent post-scaling needs connectivity
  // and stays with the caller)
  void map(const double *in, int dims, double *out, int threads = 1) const
  {
    if (!_computed)
      throw Error("The mapping has not been computed.");
    if (_constraint == Constraint::CONSERVATIVE)
      mapConservative(in, dims, out, threads);
    else
      mapConsistent(in, dims, out, threads);
  }

  void clear() // :593-600
  {
    _clusters.clear();
    _centerGrid.reset();
    _radius   = 0;
    _computed = false;
    _jit      = false;
  }

  bool                        hasComputedMapping() const { return _computed; }
  double                      clusterRadius() const { return _radius; }
  const std::vector<Cluster> &clusters() const { return _clusters; }
  int                         nIn() const { return _nIn; }  // solver-side inMesh size
  int                         nOut() const { return _nOut; } // solver-side outMesh size

private:
  Constraint           _constraint;
  int                  _dim;
  Rbf                  _f;
  Polynomial           _poly;
  unsigned             _vpc;
  double               _overlap;
  bool                 _project;
  bool                 _computed = false;
  double               _radius   = 0.0;
  int                  _nIn = 0, _nOut = 0;
  std::vector<double>  _inXYZ, _outXYZ;
  std::vector<Cluster> _clusters;
  bool                 _jit = false;
  bool                 _sampled = false; // inside computeMappingForSamples
  std::vector<int>     _samples;
  std::vector<double>  _centerXYZ;
  std::unique_ptr<PointGrid> _centerGrid;
};

// ---------------------------------------------------------------------------------------------
// Global mapping, serial branch: restates RadialBasisFctMapping (RadialBasisFctMapping.hpp:94-167,
// 199-389) with RadialBasisFctBaseMapping's dead-axis handling.
class GlobalMapping {
public:
  GlobalMapping(Constraint constraint, int dim, const Rbf &f, std::array<bool, 3> deadAxis, Polynomial polynomial)
      : _constraint(constraint), _dim(dim), _f(f), _dead(deadAxis), _poly(polynomial)
  {
    // RadialBasisFctMapping.hpp:90
    if (f.isStrictlyPositiveDefinite() && polynomial == Polynomial::ON)
      throw Error("The integrated polynomial (polynomial=\"on\") is not supported for the selected radial-basis function. Please select another radial-basis function or change the polynomial configuration.");
    // RadialBasisFctBaseMapping.hpp:64-86: 2-D meshes always have a dead z axis
    if (dim == 2)
      _dead[2] = true;
  }

  void computeMapping(const double *inputXYZ, int nInput, const double *outputXYZ, int nOutput)
  {
    const bool    cons   = _constraint == Constraint::CONSERVATIVE;
    const double *inXYZ  = cons ? outputXYZ : inputXYZ;
    const double *outXYZ = cons ? inputXYZ : outputXYZ;
    const int     n = cons ? nOutput : nInput, m = cons ? nInput : nOutput;
    std::vector<int> inIDs(n), outIDs(m);
    std::iota(inIDs.begin(), inIDs.end(), 0);
    std::iota(outIDs.begin(), outIDs.end(), 0);
    _solver   = RbfSolver(_f, inXYZ, inIDs, outXYZ, outIDs, _dead, _poly);
    _nIn      = n;
    _nOut     = m;
    _computed = true;
  }

  void map(const double *in, int dims, double *out) const
  {
    if (!_computed)
      throw Error("The mapping has not been computed.");
    if (_constraint == Constraint::CONSERVATIVE) { // :199-308 (serial branch)
      Mat u(_solver.outputSize(), dims);
      for (int i = 0; i < _nOut; ++i)
        for (int d = 0; d < dims; ++d)
          u(i, d) = in[static_cast<size_t>(i) * dims + d];
      const Mat r = _solver.solveConservative(u, _poly);
      for (int i = 0; i < _nIn; ++i) // polynomial rows stripped (:264)
        for (int d = 0; d < dims; ++d)
          out[static_cast<size_t>(i) * dims + d] = r(i, d);
    } else { // :310-389
      Mat f(_solver.inputSize(), dims); // trailing polynomial rows stay zero (:365-367)
      for (int i = 0; i < _nIn; ++i)
        for (int d = 0; d < dims; ++d)
          f(i, d) = in[static_cast<size_t>(i) * dims + d];
      const Mat r = _solver.solveConsistent(f, _poly);
      for (int i = 0; i < _nOut; ++i)
        for (int d = 0; d < dims; ++d)
          out[static_cast<size_t>(i) * dims + d] = r(i, d);
    }
  }

  void clear()
  {
    _solver   = RbfSolver();
    _computed = false;
  }
  bool             hasComputedMapping() const { return _computed; }
  const RbfSolver &solver() const { return _solver; }

private:
  Constraint          _constraint;
  int                 _dim;
  Rbf                 _f;
  std::array<bool, 3> _dead;
  Polynomial          _poly;
  RbfSolver           _solver;
  int                 _nIn = 0, _nOut = 0;
  bool                _computed = false;
};

} // namespace orc
