// rbf_oracle_capi.cpp — C entry points of the CPU ORACLE (test infrastructure, NOT the product).
// Loaded through ctypes by oracle/__init__.py; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it.
#include "rbf_oracle_mapping.hpp"

#include <chrono>

namespace {
thread_local std::string g_error;

template <typename F>
int guarded(F &&f)
{
  try {
    f();
    g_error.clear();
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return 1;
  }
}

orc::Rbf makeRbf(int basis, double param, double gaussSupport)
{
  if (!(gaussSupport > 0) && !(gaussSupport <= 0)) // NaN -> "not given"
    gaussSupport = std::numeric_limits<double>::infinity();
  return orc::Rbf::make(static_cast<orc::BasisFunction>(basis), param, gaussSupport);
}
} // namespace

extern "C" {

const char *orc_last_error() { return g_error.c_str(); }

int orc_basis_eval(int basis, double param, double gaussSupport, const double *r, int n, double *out)
{
  return guarded([&] {
    const orc::Rbf f = makeRbf(basis, param, gaussSupport);
    for (int i = 0; i < n; ++i)
      out[i] = f(r[i]);
  });
}

int orc_basis_params(int basis, double param, double gaussSupport, double *p3, int *spd, double *support)
{
  return guarded([&] {
    const orc::Rbf f = makeRbf(basis, param, gaussSupport);
    p3[0]            = f.p1;
    p3[1]            = f.p2;
    p3[2]            = f.p3;
    *spd             = f.isStrictlyPositiveDefinite();
    *support         = f.getSupportRadius();
  });
}

// ---- clustering only (PartitionOfUnityClusteringTest.cpp) ----
int orc_create_clustering(const double *inXYZ, int nIn, const double *outXYZ, int nOut, int dim, double overlap, unsigned vpc, int project,
                          double *radius, double *centers, int maxCenters, int *count)
{
  return guarded([&] {
    orc::PointGrid   gi(inXYZ, nIn), go(outXYZ, nOut);
    orc::Clustering  c = orc::createClustering(gi, go, dim, overlap, vpc, project != 0);
    *radius            = c.radius;
    *count             = c.count();
    const int n        = std::min(c.count(), maxCenters);
    if (centers)
      std::copy(c.centers.begin(), c.centers.begin() + 3 * static_cast<size_t>(n), centers);
  });
}

// impl::estimateClusterRadius (impl/CreateClustering.hpp:223-278) on an arbitrary mesh / bounding box, as
// PartitionOfUnityMapping::tagMeshFirstRound calls it (PartitionOfUnityMapping.hpp:520-521)
int orc_estimate_cluster_radius(const double *xyz, int n, int dim, unsigned vpc, const double *bbMin, const double *bbMax, double *radius)
{
  return guarded([&] {
    orc::PointGrid g(xyz, n);
    orc::Vec3      lo{{bbMin[0], bbMin[1], bbMin[2]}}, hi{{bbMax[0], bbMax[1], bbMax[2]}};
    *radius = orc::estimateClusterRadius(vpc, g, dim, lo, hi);
  });
}

// ---- partition of unity ----
void *orc_pum_create(int constraint, int dim, int basis, double param, double gaussSupport, int polynomial, unsigned vpc, double overlap, int project)
{
  void *h = nullptr;
  guarded([&] {
    h = new orc::PumMapping(static_cast<orc::Constraint>(constraint), dim, makeRbf(basis, param, gaussSupport), static_cast<orc::Polynomial>(polynomial), vpc, overlap, project != 0);
  });
  return h;
}
int orc_pum_compute(void *h, const double *inXYZ, int nIn, const double *outXYZ, int nOut, int threads)
{
  return guarded([&] { static_cast<orc::PumMapping *>(h)->computeMapping(inXYZ, nIn, outXYZ, nOut, threads); });
}
int orc_pum_compute_sampled(void *h, const double *inXYZ, int nIn, const double *outXYZ, int nOut, const int *samples, int nSamples, int threads)
{
  return guarded([&] { static_cast<orc::PumMapping *>(h)->computeMappingForSamples(inXYZ, nIn, outXYZ, nOut, samples, nSamples, threads); });
}
int orc_pum_map(void *h, const double *in, int dims, double *out, int threads)
{
  return guarded([&] { static_cast<orc::PumMapping *>(h)->map(in, dims, out, threads); });
}
// ---- just-in-time mapping (PartitionOfUnityMapping.hpp:382-476) ----
int orc_pum_compute_jit(void *h, const double *meshXYZ, int n, int threads)
{
  return guarded([&] { static_cast<orc::PumMapping *>(h)->computeMappingJustInTime(meshXYZ, n, threads); });
}
void *orc_pum_cache_create(void *h, int dims)
{
  orc::PumMapping::Cache *c = nullptr;
  guarded([&] {
    c = new orc::PumMapping::Cache();
    static_cast<orc::PumMapping *>(h)->initializeCache(*c, dims);
  });
  return c;
}
void orc_pum_cache_destroy(void *c) { delete static_cast<orc::PumMapping::Cache *>(c); }
int  orc_pum_cache_reset(void *h, void *c)
{
  return guarded([&] {
    auto *cache = static_cast<orc::PumMapping::Cache *>(c);
    static_cast<orc::PumMapping *>(h)->initializeCache(*cache, cache->dims);
  });
}
int orc_pum_cache_update(void *h, void *c, const double *in)
{
  return guarded([&] { static_cast<orc::PumMapping *>(h)->updateCache(*static_cast<orc::PumMapping::Cache *>(c), in); });
}
int orc_pum_map_consistent_at(void *h, void *c, const double *coords, int nPts, double *out)
{
  return guarded([&] { static_cast<orc::PumMapping *>(h)->mapConsistentAt(coords, nPts, *static_cast<orc::PumMapping::Cache *>(c), out); });
}
int orc_pum_map_conservative_at(void *h, void *c, const double *coords, int nPts, const double *source)
{
  return guarded([&] { static_cast<orc::PumMapping *>(h)->mapConservativeAt(coords, nPts, source, *static_cast<orc::PumMapping::Cache *>(c)); });
}
int orc_pum_complete_jit(void *h, void *c, double *out)
{
  return guarded([&] { static_cast<orc::PumMapping *>(h)->completeJustInTimeMapping(*static_cast<orc::PumMapping::Cache *>(c), out); });
}
void   orc_pum_clear(void *h) { static_cast<orc::PumMapping *>(h)->clear(); }
void   orc_pum_destroy(void *h) { delete static_cast<orc::PumMapping *>(h); }
double orc_pum_radius(void *h) { return static_cast<orc::PumMapping *>(h)->clusterRadius(); }
int    orc_pum_num_clusters(void *h) { return static_cast<int>(static_cast<orc::PumMapping *>(h)->clusters().size()); }
int    orc_pum_cluster_sizes(void *h, int c, double *center, int *nIn, int *nOut, int *polyParams)
{
  return guarded([&] {
    const auto &k = static_cast<orc::PumMapping *>(h)->clusters().at(c);
    for (int d = 0; d < 3; ++d)
      center[d] = k.center[d];
    *nIn        = static_cast<int>(k.inIDs.size());
    *nOut       = static_cast<int>(k.outIDs.size());
    *polyParams = k.solver.numberOfPolynomials();
  });
}
int orc_pum_cluster_data(void *h, int c, int *inIDs, int *outIDs, double *weights, double *choleskyL)
{
  return guarded([&] {
    const auto &k = static_cast<orc::PumMapping *>(h)->clusters().at(c);
    if (inIDs)
      std::copy(k.inIDs.begin(), k.inIDs.end(), inIDs);
    if (outIDs)
      std::copy(k.outIDs.begin(), k.outIDs.end(), outIDs);
    if (weights)
      std::copy(k.weights.begin(), k.weights.end(), weights);
    if (choleskyL) {
      const auto &L = k.solver.choleskyL();
      std::copy(L.a.begin(), L.a.end(), choleskyL);
    }
  });
}

// ---- global direct ----
void *orc_global_create(int constraint, int dim, int basis, double param, double gaussSupport, const int *deadAxis, int polynomial)
{
  void *h = nullptr;
  guarded([&] {
    h = new orc::GlobalMapping(static_cast<orc::Constraint>(constraint), dim, makeRbf(basis, param, gaussSupport),
                               {{deadAxis[0] != 0, deadAxis[1] != 0, deadAxis[2] != 0}}, static_cast<orc::Polynomial>(polynomial));
  });
  return h;
}
int orc_global_compute(void *h, const double *inXYZ, int nIn, const double *outXYZ, int nOut)
{
  return guarded([&] { static_cast<orc::GlobalMapping *>(h)->computeMapping(inXYZ, nIn, outXYZ, nOut); });
}
int orc_global_map(void *h, const double *in, int dims, double *out)
{
  return guarded([&] { static_cast<orc::GlobalMapping *>(h)->map(in, dims, out); });
}
void orc_global_clear(void *h) { static_cast<orc::GlobalMapping *>(h)->clear(); }
void orc_global_destroy(void *h) { delete static_cast<orc::GlobalMapping *>(h); }

// ---- small dense kernels, exposed so the tests can pin them against LAPACK (scipy) ----
int orc_cholesky(double *A, int n) // column-major in/out (lower), returns 0 on success
{
  int rc = 1;
  guarded([&] {
    orc::Mat M(n, n);
    std::copy(A, A + static_cast<size_t>(n) * n, M.a.begin());
    orc::Cholesky c;
    c.compute(std::move(M));
    std::copy(c.L.a.begin(), c.L.a.end(), A);
    rc = c.ok ? 0 : 1;
  });
  return rc;
}
int orc_qr_solve(const double *A, int rows, int cols, const double *B, int nrhs, double *X, int transposed, int *rank)
{
  return guarded([&] {
    orc::Mat M(rows, cols);
    std::copy(A, A + static_cast<size_t>(rows) * cols, M.a.begin());
    orc::ColPivQR qr;
    qr.compute(std::move(M));
    *rank           = qr.rank();
    const int brows = transposed ? cols : rows;
    orc::Mat  Bm(brows, nrhs);
    std::copy(B, B + static_cast<size_t>(brows) * nrhs, Bm.a.begin());
    const orc::Mat R = transposed ? qr.solveTransposed(Bm) : qr.solve(Bm);
    std::copy(R.a.begin(), R.a.end(), X);
  });
}
int orc_singular_values(const double *A, int rows, int cols, double *sv)
{
  return guarded([&] {
    orc::Mat M(rows, cols);
    std::copy(A, A + static_cast<size_t>(rows) * cols, M.a.begin());
    const auto s = orc::singularValues(std::move(M));
    std::copy(s.begin(), s.end(), sv);
  });
}

// ---- timing helper for bench.py: one computeMapping + one map on the host cores ----
// returns seconds for (compute, map) in t[0], t[1]
int orc_pum_time(int constraint, int dim, int basis, double param, int polynomial, unsigned vpc, double overlap, int project,
                 const double *inXYZ, int nIn, const double *outXYZ, int nOut, const double *in, int dims, double *out, int threads, double *t,
                 int *nClusters)
{
  return guarded([&] {
    orc::PumMapping m(static_cast<orc::Constraint>(constraint), dim, makeRbf(basis, param, std::numeric_limits<double>::infinity()),
                      static_cast<orc::Polynomial>(polynomial), vpc, overlap, project != 0);
    const auto t0 = std::chrono::steady_clock::now();
    m.computeMapping(inXYZ, nIn, outXYZ, nOut, threads);
    const auto t1 = std::chrono::steady_clock::now();
    m.map(in, dims, out, threads);
    const auto t2 = std::chrono::steady_clock::now();
    t[0]          = std::chrono::duration<double>(t1 - t0).count();
    t[1]          = std::chrono::duration<double>(t2 - t1).count();
    *nClusters    = static_cast<int>(m.clusters().size());
  });
}

// ---- CPU-N baseline: `ranks` independent mappings, one per host thread, like one preCICE rank per core (each rank
// clusters and maps its own partition without communication, BatchedRBFSolver.hpp:29-31). Partition r uses the
// vertices [inOff[r], inOff[r+1]) / [outOff[r], outOff[r+1]). t[0] = wall seconds of the slowest rank for
// computeMapping, t[1] for map.
int orc_pum_time_ranks(int constraint, int dim, int basis, double param, int polynomial, unsigned vpc, double overlap, int project,
                       const double *inXYZ, const long long *inOff, const double *outXYZ, const long long *outOff, const double *in, int dims,
                       double *out, int ranks, double *t, long long *nClusters)
{
  return guarded([&] {
    std::vector<double>      tc(ranks, 0.0), tm(ranks, 0.0);
    std::vector<long long>   ncl(ranks, 0);
    std::vector<std::string> errors(ranks);
    std::vector<std::thread> pool;
    for (int r = 0; r < ranks; ++r)
      pool.emplace_back([&, r]() {
        try {
          orc::PumMapping m(static_cast<orc::Constraint>(constraint), dim, makeRbf(basis, param, std::numeric_limits<double>::infinity()),
                            static_cast<orc::Polynomial>(polynomial), vpc, overlap, project != 0);
          const int  nIn = static_cast<int>(inOff[r + 1] - inOff[r]), nOut = static_cast<int>(outOff[r + 1] - outOff[r]);
          const auto t0 = std::chrono::steady_clock::now();
          m.computeMapping(inXYZ + 3 * inOff[r], nIn, outXYZ + 3 * outOff[r], nOut, 1);
          const auto t1 = std::chrono::steady_clock::now();
          m.map(in + inOff[r] * dims, dims, out + outOff[r] * dims, 1);
          const auto t2 = std::chrono::steady_clock::now();
          tc[r]         = std::chrono::duration<double>(t1 - t0).count();
          tm[r]         = std::chrono::duration<double>(t2 - t1).count();
          ncl[r]        = static_cast<long long>(m.clusters().size());
        } catch (const std::exception &e) {
          errors[r] = e.what();
        }
      });
    for (auto &th : pool)
      th.join();
    for (auto &e : errors)
      if (!e.empty())
        throw orc::Error(e);
    t[0]       = *std::max_element(tc.begin(), tc.end());
    t[1]       = *std::max_element(tm.begin(), tm.end());
    *nClusters = std::accumulate(ncl.begin(), ncl.end(), 0LL);
  });
}

} // extern "C"
