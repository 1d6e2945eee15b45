// pum.cu — host driver of the partition-of-unity mapping (rbf-pum-direct).
//
// Mirrors mapping::PartitionOfUnityMapping (mapping/PartitionOfUnityMapping.hpp:181-380) and
// impl::createClustering / estimateClusterRadius (mapping/impl/CreateClustering.hpp:223-505). The scalar
// lattice arithmetic stays on the host (a few dozen flops); everything that touches vertices runs in kernels.
#include "mapping.cuh"
#include "pum.cuh"

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <functional>
#include <limits>
#include <numeric>
#include <sstream>

namespace rbfmap {
namespace {

class PumMapping final : public DeviceMapping {
public:
  explicit PumMapping(const rbfmap_config &cfg)
      : DeviceMapping(cfg)
  {
    // PartitionOfUnityMapping.hpp:163-168
    if (cfg.polynomial == RBFMAP_POLYNOMIAL_ON)
      throw MapError(RBFMAP_ERR_INVALID, "Integrated polynomial is not supported for partition of unity data mappings.");
    if (!(cfg.relative_overlap < 1.0))
      throw MapError(RBFMAP_ERR_INVALID, "The relative overlap has to be smaller than one.");
    if (cfg.vertices_per_cluster == 0)
      throw MapError(RBFMAP_ERR_INVALID, "The number of vertices per cluster has to be greater zero.");
    // thin-plate splines, multiquadrics, volume splines: the reference factorises with ColPivHouseholderQR
    // (RadialBasisFctSolver.hpp:354-358); here the explicit inverse of every cluster matrix (pum_inverse.cu)
    _inverse = !basisIsSpd(cfg.basis);
  }

  std::string name() const override { return "partition-of-unity RBF (cuda-executor)"; }

  // one mapping sharded over the ranks of the group, see include/rbfmap.h
  void distributedInit(const char *uniqueId, int rank, int world) override
  {
    if (_computed)
      throw MapError(RBFMAP_ERR_STATE, "Please clear the mapping before joining a group."); // the shards depend on the group
    if (world > 1 && !uniqueId && (_cfg.shard_mode == RBFMAP_SHARD_EXCHANGE || isConservative()))
      throw MapError(RBFMAP_ERR_INVALID, "a group without NCCL id can only shard consistent mappings in duplicate mode (no collective).");
    _group.init(uniqueId, rank, world);
  }

  void computeMapping(const double *dInputXyz, int64_t nInput, const double *dOutputXyz, int64_t nOutput) override
  {
    computeImpl(dInputXyz, nInput, dOutputXyz, nOutput, cudaMemcpyDeviceToDevice);
  }
  bool computeMappingHost(const double *hInputXyz, int64_t nInput, const double *hOutputXyz, int64_t nOutput) override
  {
    computeImpl(hInputXyz, nInput, hOutputXyz, nOutput, cudaMemcpyHostToDevice);
    return true;
  }
  // ---- just-in-time mapping ----
  void jitComputeMapping(const double *hMeshXyz, int64_t n) override
  {
    if (_group.active())
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "just-in-time mapping is not available on a sharded mapping.");
    if (_inverse)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "just-in-time mapping requires a strictly positive definite basis function.");
    _jit = true;
    try {
      if (isConservative())
        computeImpl(nullptr, 0, hMeshXyz, n, cudaMemcpyHostToDevice);
      else
        computeImpl(hMeshXyz, n, nullptr, 0, cudaMemcpyHostToDevice);
    } catch (...) {
      _jit = false;
      throw;
    }
  }
  std::unique_ptr<JitCache> jitCreateCache(int dims) override
  {
    requireJit();
    if (dims < 1)
      throw MapError(RBFMAP_ERR_INVALID, "data dimension has to be positive");
    auto c  = std::make_unique<JitCache>();
    c->dims = dims;
    c->coef.reserve(static_cast<size_t>(std::max<int64_t>(_mem.sumIn, 1)) * dims);
    c->poly.reserve(static_cast<size_t>(std::max<int64_t>(_nClusters, 1)) * 4 * dims);
    jitResetCache(*c);
    return c;
  }
  void jitResetCache(JitCache &c) override
  {
    requireJit();
    checkCache(c);
    RBF_CUDA(cudaMemsetAsync(c.coef.p, 0, c.coef.bytes(), _stream));
    RBF_CUDA(cudaMemsetAsync(c.poly.p, 0, c.poly.bytes(), _stream));
    RBF_CUDA(cudaStreamSynchronize(_stream));
  }
  void jitUpdateCache(JitCache &c, const double *hValues) override;
  void jitMapConsistentAt(JitCache &c, const double *hCoords, int64_t nPts, double *hOut) override;
  void jitMapConservativeAt(JitCache &c, const double *hCoords, int64_t nPts, const double *hSource) override;
  void jitComplete(JitCache &c, double *hOut) override;

  ~PumMapping() override
  {
    if (_outReady)
      cudaEventDestroy(_outReady);
    if (_allocReady)
      cudaEventDestroy(_allocReady);
    if (_copyStream)
      cudaStreamDestroy(_copyStream);
  }
  double estimateClusterRadiusHost(const double *hXyz, int64_t n, const double *bbMin, const double *bbMax) override
  {
    AllocStreamScope allocScope(_stream);
    DevBuf<double>   d;
    d.alloc(3 * n);
    RBF_CUDA(cudaMemcpyAsync(d.p, hXyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, _stream));
    return estimateClusterRadius(d.p, n, bbMin, bbMax);
  }
  // PartitionOfUnityMapping.hpp:479-533
  void tagMeshFirstRound(const double *hFilterXyz, int64_t nFilter, const double *hOtherXyz, int64_t nOther, unsigned char *hTags) override
  {
    if (nOther == 0 || nFilter == 0)
      return; // :491-492
    AllocStreamScope allocScope(_stream);
    DevBuf<double>   dFilter, dOther;
    dFilter.alloc(3 * nFilter);
    dOther.alloc(3 * nOther);
    RBF_CUDA(cudaMemcpyAsync(dFilter.p, hFilterXyz, 3 * nFilter * sizeof(double), cudaMemcpyHostToDevice, _stream));
    RBF_CUDA(cudaMemcpyAsync(dOther.p, hOtherXyz, 3 * nOther * sizeof(double), cudaMemcpyHostToDevice, _stream));
    double lo[3], hi[3];
    computeBounds(dOther.p, nOther, lo, hi, _stream); // :513 the local bounding box
    if (_radius == 0) // :520-521: estimated on the unfiltered remote mesh, kept until clear()
      _radius = estimateClusterRadius(dFilter.p, nFilter, lo, hi);
    tagInsideBox(dFilter.p, nFilter, lo, hi, 2 * _radius, hTags); // :527-532
  }
  void tagMeshSecondRound(const double *, int64_t, const unsigned char *, unsigned char *) override {} // :535-539
  void mapConsistent(const double *dIn, int dims, double *dOut) override { map(false, dIn, dims, dOut); }
  void mapConservative(const double *dIn, int dims, double *dOut) override { map(true, dIn, dims, dOut); }
  void clear() override
  {
    // PartitionOfUnityMapping.hpp:593-600
    _inXyz.release();
    _outXyz.release();
    _inPtr = _outPtr = nullptr;
    _centers.release();
    _mem = PumMembership();
    _poly.release();
    _wsum.release();
    _outRec.release();
    _inRec.release();
    _wcount.release();
    _buckets.order.release();
    _buckets.entries.release();
    _centerGrid = BinGrid();
    _globalIdx.release();
    _evalOff.release();
    _evalReady  = false;
    _blendIndex = PumVertexIndex();
    _jit        = false;
    _nClusters  = 0;
    _nClustersGlobal = 0;
    _radius    = 0.0;
    _computed  = false;
    stats      = rbfmap_stats{};
    stats.shard_axis = -1;
  }

  // inspection (tests)
  int64_t              nClusters() const { return _nClusters; }
  const DevBuf<double> &centers() const { return _centers; }
  const PumMembership  &membership() const { return _mem; }

private:
  void requireJit() const
  {
    if (!_computed || !_jit)
      throw MapError(RBFMAP_ERR_STATE, "The mapping has not been computed as a just-in-time mapping; call computeMappingJustInTime() first.");
  }
  void checkCache(const JitCache &c) const
  {
    if (c.coef.n < static_cast<size_t>(std::max<int64_t>(_mem.sumIn, 1)) * c.dims || c.poly.n < static_cast<size_t>(std::max<int64_t>(_nClusters, 1)) * 4 * c.dims)
      throw MapError(RBFMAP_ERR_STATE, "The data cache does not belong to the current mapping; create it after computeMappingJustInTime().");
  }
  void throwUnassigned(const double *dCoords, int64_t bad, bool cons) const;
  void   computeImpl(const double *srcInputXyz, int64_t nInput, const double *srcOutputXyz, int64_t nOutput, cudaMemcpyKind kind);
  void   map(bool conservative, const double *dIn, int dims, double *dOut);
  double estimateClusterRadius(const double *dXyz, int64_t n, const double bbMin[3], const double bbMax[3]);
  PumSolveArgs solveArgs() const;
  // sharded mapping: replaces the complete centre list by the centres this rank processes
  // sharded mapping: chooseSlab before the clustering (cuts, region), selectClusters after it
  void chooseSlab(const double bbMin[3], const double bbMax[3], const int latticeCount[3], bool haveLattice);
  void selectClusters();
  bool exchangeMode() const { return _group.active() && (_cfg.shard_mode == RBFMAP_SHARD_EXCHANGE || isConservative()); }
  bool duplicateMode() const { return _group.active() && !exchangeMode(); }
  bool mapResultIsPartial() const override { return duplicateMode(); }

  DevBuf<double>  _inXyz, _outXyz; // solver-side inMesh / outMesh (owned copies: host-pointer and just-in-time mappings)
  // what computeMapping reads: the owned copies, or - device-pointer entry - the caller's buffers themselves (the call is
  // synchronous and nothing refers to the meshes afterwards: the map kernels read the cluster-ordered records), which
  // saves a 480 MB device-to-device copy per computeMapping at 10M; null outside computeMapping in that case
  const double *_inPtr = nullptr, *_outPtr = nullptr;
  int64_t         _nIn = 0, _nOut = 0;
  double          _radius    = 0.0;
  int64_t         _nClusters = 0; // clusters this rank processes (all of them unless the mapping is sharded)
  int64_t         _nClustersGlobal = 0;
  int64_t         _ownedClusters   = 0; // sharded mapping: clusters whose centre lies in this rank's slab
  PumSlab         _slab{};        // sharded mapping: what this rank owns
  DevBuf<int>     _globalIdx;     // sharded mapping: index of each local cluster in the complete list
  DevBuf<double>  _centers;
  PumMembership   _mem;
  PumScratch      _scratch; // survives clear(), see pum.cuh
  ScratchBuf<double> _mat; // packed factors: grow-only, survives clear() like _scratch (a remeshing step changes its size every
                           // time; through the stream-ordered pool that cost 0.8 - 4 ms per step for a 2 GB buffer)
  DevBuf<PolyRec> _poly;
  DevBuf<double>  _wsum;
  DevBuf<double4> _outRec;
  DevBuf<double>  _inRec;
  DevBuf<int>     _wcount;
  PumBuckets      _buckets;
  // execution-mode minimal-compute (computeEvaluationOffline, PartitionOfUnityMapping.hpp:57): stored evaluation matrices
  ScratchBuf<double> _eval; // grow-only like _mat
  DevBuf<int64_t>    _evalOff;
  bool               _evalReady = false;
  // deterministic blend (rbfmap_config.deterministic)
  PumVertexIndex  _blendIndex;
  bool            _inverse = false; // basis function not strictly positive definite: explicit inverses instead of factors
  bool            _jit = false; // the other mesh is a just-in-time mesh
  BinGrid         _centerGrid;  // cluster centres, kept for just-in-time queries (PartitionOfUnityMapping.hpp:285-287)
  // upload of the solver-side outMesh beside the work on the inMesh (computeMappingHost)
  cudaStream_t _copyStream = nullptr;
  cudaEvent_t  _allocReady = nullptr, _outReady = nullptr;
};

PumSolveArgs PumMapping::solveArgs() const
{
  PumSolveArgs a{};
  a.rbf     = _rbf;
  a.dim     = _cfg.dimensions;
  a.radius  = _radius;
  a.inXyz   = _inPtr;
  a.outXyz  = _outPtr;
  a.centers = _centers.p;
  a.nIn     = _mem.nIn.p;
  a.nOut    = _mem.nOut.p;
  a.inOff   = reinterpret_cast<const long long *>(_mem.inOff.p);
  a.outOff  = reinterpret_cast<const long long *>(_mem.outOff.p);
  a.matOff  = reinterpret_cast<const long long *>(_mem.matOff.p);
  a.inIds   = _mem.inIds.p;
  a.outIds  = _mem.outIds.p;
  a.mat     = _mat.p;
  a.poly    = _poly.p;
  a.wsum    = _wsum.p;
  a.wcount  = _wcount.p;
  a.outRec  = _outRec.p;
  a.inRec   = _inRec.p;
  a.inverse = _inverse;
  if (_evalReady) {
    a.eval    = _eval.p;
    a.evalOff = reinterpret_cast<const long long *>(_evalOff.p);
  }
  return a;
}

// impl/CreateClustering.hpp:223-278
double PumMapping::estimateClusterRadius(const double *dXyz, int64_t nXyz, const double bbMin[3], const double bbMax[3])
{
  const int dim = _cfg.dimensions;
  double    center[3] = {0, 0, 0};
  for (int d = 0; d < dim; ++d)
    center[d] = (bbMax[d] + bbMin[d]) * 0.5;
  std::vector<double> probes(center, center + 3);
  for (int d = 0; d < dim; ++d) {
    double s[3] = {center[0], center[1], center[2]};
    const double edge = std::abs(bbMax[d] - bbMin[d]);
    s[d] -= edge * 0.25;
    probes.insert(probes.end(), s, s + 3);
    s[d] += edge * 0.5;
    probes.insert(probes.end(), s, s + 3);
  }
  const int        nProbes = static_cast<int>(probes.size() / 3);
  std::vector<int> samples(nProbes);
  std::vector<double> sampleXyz(3 * static_cast<size_t>(nProbes));
  nearestVertices(dXyz, nXyz, probes.data(), nProbes, samples.data(), _stream, sampleXyz.data());

  // starting radius for the k-nearest search from the mean density of the bounding box
  double vol  = 1.0;
  int    live = 0;
  for (int d = 0; d < dim; ++d)
    if (bbMax[d] > bbMin[d]) {
      vol *= (bbMax[d] - bbMin[d]);
      ++live;
    }
  const double guess = live > 0 ? 0.8 * std::pow(vol * static_cast<double>(_cfg.vertices_per_cluster) / static_cast<double>(nXyz), 1.0 / live) : 1.0;
  std::vector<double> radii(nProbes);
  kthNeighbourDistance(dXyz, nXyz, samples.data(), nProbes, static_cast<int>(_cfg.vertices_per_cluster), guess, radii.data(), _stream, sampleXyz.data());
  const size_t middle = radii.size() / 2;
  std::nth_element(radii.begin(), radii.begin() + middle, radii.end());
  return radii[middle];
}

// SURVEY.md 8(e), include/rbfmap.h: slabs along the lattice axis with the most layers. The cuts are chosen BEFORE the
// clustering, from the histogram of the in-vertices along that axis (clusters follow the vertex density), so that every
// rank can restrict binning and clustering to its own region of the lattice; every rank computes the same cuts.
void PumMapping::chooseSlab(const double bbMin[3], const double bbMax[3], const int latticeCount[3], bool haveLattice)
{
  const int    world = _group.world, rank = _group.rank;
  const double inf   = std::numeric_limits<double>::infinity();
  int          axis  = 0;
  for (int d = 1; d < _cfg.dimensions; ++d)
    if (latticeCount[d] > latticeCount[axis])
      axis = d;
  _slab.axis = axis;
  double lo = -inf, hi = inf;
  const double edge = bbMax[axis] - bbMin[axis];
  if (!haveLattice || !(edge > 0.0)) {
    // single-cluster fallback / degenerate axis: rank 0 takes everything
    if (rank != 0)
      lo = hi = inf;
  } else {
    constexpr int kBins    = 4096;
    const double  binWidth = edge / kBins;
    const auto    hist     = pumVertexHistogram(_inPtr, _nIn, axis, bbMin[axis], binWidth, kBins, _stream);
    // duplicate mode: a rank also processes the clusters within one radius of its slab
    const int        haloBins = exchangeMode() ? 0 : static_cast<int>(std::ceil(_radius / binWidth)) + 1;
    std::vector<int> cuts(world + 1);
    pumPartitionSlabs(hist.data(), kBins, world, haloBins, cuts.data());
    if (rank > 0)
      lo = bbMin[axis] + cuts[rank] * binWidth;
    if (rank < world - 1)
      hi = bbMin[axis] + cuts[rank + 1] * binWidth;
    if (rank > 0 && cuts[rank] >= kBins)
      lo = hi = inf; // more ranks than occupied bins
    else if (!(lo < hi))
      lo = hi = inf;
  }
  _slab.lo = lo;
  _slab.hi = hi;
  stats.shard_axis = axis;
  stats.shard_lo   = lo;
  stats.shard_hi   = hi;
}

// After the (region-restricted) clustering: keeps the centres this rank processes - its own (exchange mode), or every centre
// whose sphere reaches into the slab (duplicate mode; |v - c| < r implies |v_a - c_a| < r, the margin covers the rounding
// of the comparison) - and counts the clusters of the whole mapping (every centre is owned by exactly one slab).
void PumMapping::selectClusters()
{
  const double inf    = std::numeric_limits<double>::infinity();
  const double margin = exchangeMode() ? 0.0 : _radius * (1.0 + 1e-12) + 1e-300;
  double       selLo = _slab.lo - margin, selHi = _slab.hi + margin;
  if (_slab.lo == inf) // empty slab
    selLo = selHi = inf;
  const int64_t  nRegion = _nClusters;
  DevBuf<double> local, owned;
  DevBuf<int>    ownedIdx;
  const int64_t  nOwned = (margin > 0.0 && _group.connected()) ? pumSelectSlab(_centers.p, nRegion, _slab.axis, _slab.lo, _slab.hi, owned, ownedIdx, _stream) : -1;
  _nClusters            = pumSelectSlab(_centers.p, nRegion, _slab.axis, selLo, selHi, local, _globalIdx, _stream);
  _centers              = std::move(local);
  // clusters of the whole mapping = sum of the owned ones over the ranks: agreed on at the end of computeMapping together
  // with the error flags (one collective); unknown (-1) without communicator
  _ownedClusters   = nOwned >= 0 ? nOwned : _nClusters;
  _nClustersGlobal = -1;
}

void PumMapping::computeImpl(const double *dInputXyz, int64_t nInput, const double *dOutputXyz, int64_t nOutput, cudaMemcpyKind kind)
{
  if (_computed)
    throw MapError(RBFMAP_ERR_STATE, "Please clear the mapping before recomputing."); // PartitionOfUnityMapping.hpp:188
  g_launches = 0;
  AllocStreamScope allocScope(_stream);
  EventTimer       whole(_stream);
  whole.start();
  PhaseTrace trace("pum.computeMapping", _stream);
  _nInput  = nInput;
  _nOutput = nOutput;
  // PartitionOfUnityMapping.hpp:190-198
  const bool    cons   = isConservative();
  const double *srcIn  = cons ? dOutputXyz : dInputXyz;
  const double *srcOut = cons ? dInputXyz : dOutputXyz;
  _nIn                 = cons ? nOutput : nInput;
  _nOut                = cons ? nInput : nOutput;
  if (_nIn > INT_MAX / 4 || _nOut > INT_MAX / 4)
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "mesh too large for 32-bit vertex IDs");
  EventTimer replicated(_stream);
  replicated.start();
  const bool borrow = kind == cudaMemcpyDeviceToDevice && !_jit; // device-resident caller: read its buffers in place
  struct Unborrow {
    PumMapping *m;
    bool        on;
    ~Unborrow()
    {
      if (on)
        m->_inPtr = m->_outPtr = nullptr;
    }
  } unborrow{this, borrow};
  if (borrow) {
    _inXyz.release();
    _outXyz.release();
    _inPtr  = srcIn;
    _outPtr = srcOut;
  } else {
    // sharded mapping: every rank uploads 1/world of each mesh over PCIe, the slices are all-gathered over NVLink
    _inXyz.alloc(std::max<int64_t>(_group.padded(3 * _nIn), 1));
    _outXyz.alloc(std::max<int64_t>(_group.padded(3 * _nOut), 1));
    _inPtr  = _inXyz.p;
    _outPtr = _outXyz.p;
  }
  trace.phase("alloc meshes");
  if (_nIn > 0 && !borrow) {
    if (kind == cudaMemcpyDeviceToDevice)
      deviceCopy(_inXyz.p, srcIn, 3 * _nIn, _stream);
    else
      uploadReplicated(_inXyz.p, srcIn, 3 * _nIn, _stream);
  }
  // From host memory the outMesh travels on a second stream: bounds, cluster-radius estimate and binning of the inMesh
  // run while it is still in flight; the main stream waits for it (waitForOutMesh) right before its first use.
  struct PendingCopy { // the host buffer must not be in use any more when this call returns, however it returns
    cudaStream_t s       = nullptr;
    bool         pending = false;
    ~PendingCopy()
    {
      if (pending)
        cudaStreamSynchronize(s);
    }
  } outCopy;
  if (_nOut > 0 && !borrow) {
    if (kind == cudaMemcpyHostToDevice) {
      if (!_copyStream) {
        RBF_CUDA(cudaStreamCreateWithFlags(&_copyStream, cudaStreamNonBlocking));
        RBF_CUDA(cudaEventCreateWithFlags(&_allocReady, cudaEventDisableTiming));
        RBF_CUDA(cudaEventCreateWithFlags(&_outReady, cudaEventDisableTiming));
      }
      RBF_CUDA(cudaEventRecord(_allocReady, _stream)); // the buffer comes from the stream-ordered pool of _stream
      RBF_CUDA(cudaStreamWaitEvent(_copyStream, _allocReady, 0));
      if (_group.connected()) { // this rank's slice only; the all-gather is issued on the main stream in waitForOutMesh
        const int64_t slice = _group.sliceOf(3 * _nOut);
        const int64_t b = std::min(3 * _nOut, slice * _group.rank), e = std::min(3 * _nOut, b + slice);
        if (e > b)
          RBF_CUDA(cudaMemcpyAsync(_outXyz.p + b, srcOut + b, static_cast<size_t>(e - b) * sizeof(double), kind, _copyStream));
      } else {
        RBF_CUDA(cudaMemcpyAsync(_outXyz.p, srcOut, 3 * _nOut * sizeof(double), kind, _copyStream));
      }
      RBF_CUDA(cudaEventRecord(_outReady, _copyStream));
      outCopy.s       = _copyStream;
      outCopy.pending = true;
    } else if (kind == cudaMemcpyDeviceToDevice) {
      deviceCopy(_outXyz.p, srcOut, 3 * _nOut, _stream);
    } else {
      uploadReplicated(_outXyz.p, srcOut, 3 * _nOut, _stream);
    }
  }
  auto waitForOutMesh = [&] {
    if (outCopy.pending) {
      RBF_CUDA(cudaStreamWaitEvent(_stream, _outReady, 0));
      outCopy.pending = false; // ordered behind _stream from here on; every exit path below synchronises _stream
      if (_group.connected())
        _group.allGatherInPlace(_outXyz.p, _group.sliceOf(3 * _nOut), _stream);
    }
  };

  stats            = rbfmap_stats{};
  stats.shard_axis = -1;
  stats.n_in       = _nIn;
  stats.n_out      = _nOut;
  _nClusters       = 0;
  _nClustersGlobal = 0;
  _radius          = 0.0;

  // impl/CreateClustering.hpp:314-316: nothing to do for empty meshes
  const bool jit = _jit; // CreateClustering.hpp:313: a just-in-time outMesh is empty by construction
  if (_nIn == 0 || (_nOut == 0 && !jit)) {
    _computed = true;
    return;
  }
  const int dim = _cfg.dimensions;
  double    bbMin[3], bbMax[3], obMin[3], obMax[3];
  trace.phase("alloc + copy");
  computeBounds(_inPtr, _nIn, bbMin, bbMax, _stream); // :334 (bounds of the inMesh)
  trace.phase("bounds");
  auto edge = [&](int d) { return std::abs(bbMax[d] - bbMin[d]); };

  BinGrid inGrid, outGrid;
  int     latticeCount[3] = {1, 1, 1};
  bool    haveLattice = false, replicatedDone = false;
  if (static_cast<uint64_t>(_nIn) < static_cast<uint64_t>(_cfg.vertices_per_cluster) * 2) {
    // single-cluster fallback :352-359
    double cRadius = 0.0;
    for (int d = 0; d < dim; ++d)
      cRadius = d == 0 ? (bbMax[d] - bbMin[d]) : std::max(cRadius, bbMax[d] - bbMin[d]);
    if (cRadius == 0)
      cRadius = 1;
    _radius     = cRadius * 2;
    double c[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d)
      c[d] = (bbMax[d] + bbMin[d]) * 0.5;
    _centers.alloc(3);
    RBF_CUDA(cudaMemcpyAsync(_centers.p, c, sizeof(c), cudaMemcpyHostToDevice, _stream));
    RBF_CUDA(cudaStreamSynchronize(_stream));
    _nClusters = 1;
    inGrid.build(_inPtr, _nIn, bbMin, bbMax, 0.5 * _radius, _stream); // half-radius cells, 5x5x5 neighbourhoods (forCandidatesWarp<2>)
    waitForOutMesh();
    if (_nOut > 0)
      computeBounds(_outPtr, _nOut, obMin, obMax, _stream);
    else
      for (int d = 0; d < 3; ++d)
        obMin[d] = bbMin[d], obMax[d] = bbMax[d];
    outGrid.build(_outPtr, _nOut, obMin, obMax, 0.5 * _radius, _stream);
  } else {
    _radius = estimateClusterRadius(_inPtr, _nIn, bbMin, bbMax);
    trace.phase("estimateClusterRadius");
    if (!(_radius > 0.0))
      throw MapError(RBFMAP_ERR_SINGULAR, "rbf-pum-direct: the estimated cluster radius is zero (duplicated vertices?).");
    // :366-427 lattice of tentative centres
    const bool   startGridAtEdge       = _cfg.project_to_input != 0;
    const double maximumCenterDistance = std::sqrt(4. / dim) * _radius * (1 - _cfg.relative_overlap);
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
        if (edge(d) > kZeroDiff)
          nGlobal[d] += 1;
    } else {
      for (int d = 0; d < dim; ++d)
        start[d] = start[d] + 0.5 * distances[d];
    }
    std::array<unsigned, 3> nLocal{{1, 1, 1}};
    for (int d = 0; d < dim; ++d)
      if (distances[d] > 0) {
        start[d] += std::ceil(((bbMin[d] - start[d]) / distances[d]) - kZeroDiff) * distances[d];
        nLocal[d] = static_cast<unsigned>(1 + std::floor(((bbMax[d] - start[d]) / distances[d]) + kZeroDiff));
      }
    const double totalD = static_cast<double>(nLocal[0]) * nLocal[1] * nLocal[2];
    if (totalD > 1.0e9)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: the tentative cluster lattice is too large.");
    PumClustering cl;
    cl.radius = _radius;
    // :438-463 — coordinates grow by repeated addition, exactly like the reference loops
    std::vector<double> axis[3];
    for (int d = 0; d < 3; ++d) {
      axis[d].assign(nLocal[d], 0.0);
      if (d < dim) {
        double v = start[d];
        for (unsigned i = 0; i < nLocal[d]; ++i, v += distances[d])
          axis[d][i] = v;
      }
    }
    DevBuf<double> dAxis[3];
    LatticeAxes    axes{};
    for (int d = 0; d < 3; ++d) {
      dAxis[d].alloc(axis[d].size());
      RBF_CUDA(cudaMemcpyAsync(dAxis[d].p, axis[d].data(), axis[d].size() * sizeof(double), cudaMemcpyHostToDevice, _stream));
      axes.a[d]       = dAxis[d].p;
      axes.n[d]       = static_cast<int>(nLocal[d]);
      axes.start[d]   = d < dim ? start[d] : 0.0;
      axes.invDist[d] = (d < dim && distances[d] > 0) ? 1.0 / distances[d] : 0.0;
    }
    for (int d = 0; d < 3; ++d)
      latticeCount[d] = static_cast<int>(nLocal[d]);
    haveLattice = true;
    // ---- up to here every rank of a sharded mapping repeats the same work (bounds, radius estimate) ----
    stats.ms_replicated = replicated.stop();
    replicatedDone      = true;
    // Sharded mapping: this rank only clusters its own region. A final centre p that it needs lies within one radius of
    // the slab [lo, hi); p is the in-vertex nearest to a lattice point l with |p - l| < r, so l_a lies within 2 r; the
    // duplicate rule of l looks at its index neighbours, one lattice spacing s further; and the emptiness test and
    // the projection of those need the vertices within another radius. Hence: lattice layers within 2 r + s, vertices
    // within 3 r + s of the slab (plus a rounding margin). Inside that region every decision is made from complete
    // information and is identical to the one of the unsharded mapping.
    int    regionAxis = -1;
    double regionLo = 0.0, regionHi = 0.0;
    if (_group.active()) {
      chooseSlab(bbMin, bbMax, latticeCount, true);
      const int    ax = _slab.axis;
      const double sp = distances[ax], eps = 1e-9 * _radius;
      const double layerLo = _slab.lo - (2.0 * _radius + sp + eps), layerHi = _slab.hi + (2.0 * _radius + sp + eps);
      regionAxis = ax;
      regionLo   = _slab.lo - (3.0 * _radius + sp + 2.0 * eps);
      regionHi   = _slab.hi + (3.0 * _radius + sp + 2.0 * eps);
      int first = static_cast<int>(nLocal[ax]), last = -1;
      for (int i = 0; i < static_cast<int>(nLocal[ax]); ++i)
        if (axis[ax][i] >= layerLo && axis[ax][i] <= layerHi) {
          first = std::min(first, i);
          last  = std::max(last, i);
        }
      axes.rangeAxis = ax;
      axes.rangeLo   = first;
      axes.rangeHi   = last; // first > last: an empty slab
      if (!(_slab.lo < _slab.hi))
        regionLo = regionHi = std::numeric_limits<double>::quiet_NaN(); // nothing is binned
      trace.phase("slab");
    }
    // the lattice itself is never materialised: one tag per centre, one "still empty" bit per mesh that has to
    // populate the cluster (tagEmptyLattice clears them); coordinates only for the centres that survive
    cl.initLattice(axes, jit ? 1 : 3, _stream);
    trace.phase("lattice");
    // half-radius cells, 5x5x5 neighbourhoods (forCandidatesWarp<2>)
    inGrid.build(_inPtr, _nIn, bbMin, bbMax, 0.5 * _radius, _stream, regionAxis, regionLo, regionHi);
    waitForOutMesh();
    if (_nOut > 0)
      computeBounds(_outPtr, _nOut, obMin, obMax, _stream);
    else
      for (int d = 0; d < 3; ++d)
        obMin[d] = bbMin[d], obMax[d] = bbMax[d];
    outGrid.build(_outPtr, _nOut, obMin, obMax, 0.5 * _radius, _stream, regionAxis, regionLo, regionHi);
    trace.phase("bin meshes");

    // :468-471
    cl.tagEmptyLattice(inGrid, 1, _stream);
    if (!jit) // CreateClustering.hpp:469
      cl.tagEmptyLattice(outGrid, 2, _stream);
    cl.gatherLive(_stream);
    trace.phase("tagEmpty x2 + live list");
    if (_cfg.project_to_input) { // :475-496
      cl.project(inGrid, _stream);
      trace.phase("project");
      const double     threshold = 0.4 * (*std::min_element(distances.begin(), distances.end()));
      NeighbourOffsets offs{};
      const bool       twoD = nLocal[2] == 1;
      for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dz = (twoD ? 0 : -1); dz <= (twoD ? 0 : 1); ++dz) {
            if (dx == 0 && dy == 0 && dz == 0)
              continue;
            offs.off[offs.n++] = (static_cast<long long>(dx) * nLocal[1] + dy) * static_cast<long long>(nLocal[2]) + dz;
          }
      cl.tagDuplicates(offs, threshold, _stream);
      trace.phase("tagDuplicates");
      if (!jit) // CreateClustering.hpp:492
        cl.tagEmpty(outGrid, _stream);
      trace.phase("tagEmpty");
    }
    _nClusters = cl.compact(_centers, _stream);
    trace.phase("compact");
    if (_nClusters == 0 && !_group.active())
      throw MapError(RBFMAP_ERR_EMPTY, "Too many vertices have been filtered out."); // :502
  }
  _nClustersGlobal = _nClusters;
  if (!replicatedDone)
    stats.ms_replicated = replicated.stop();
  if (_group.active()) {
    if (!haveLattice)
      chooseSlab(bbMin, bbMax, latticeCount, false); // single-cluster fallback: rank 0 takes it
    selectClusters();
    trace.phase("select clusters");
  }
  // Sharded mapping: the three things the ranks have to agree on - lowest unassigned output vertex, lowest cluster with a
  // singular matrix, number of clusters - travel in ONE pair of grouped all-reduces after the factorisation, so that
  // every rank raises the same error (a rank that stopped early would leave the others waiting in the next collective).
  DevBuf<int> agree; // [0] first unassigned vertex (min), [1] failing cluster (min), [2] -(largest cluster) (min), [3] owned clusters (sum)
  if (_group.active()) {
    agree.alloc(4);
    const int init[4] = {0x7f7f7f7f, INT_MAX, 0, static_cast<int>(_ownedClusters)};
    RBF_CUDA(cudaMemcpyAsync(agree.p, init, sizeof(init), cudaMemcpyHostToDevice, _stream)); // pageable source: staged before the call returns
  }
  const PumSlab *outSlab = duplicateMode() ? &_slab : nullptr;

  stats.n_clusters        = _nClusters;
  stats.n_clusters_global = _nClustersGlobal;
  stats.cluster_radius    = _radius;

  // per-cluster vertex sets (SphericalVertexCluster.hpp:153-162)
  _wsum.alloc(_nOut);
  _wcount.alloc(_nOut);
  if (_nOut > 0) {
    RBF_CUDA(cudaMemsetAsync(_wsum.p, 0, _nOut * sizeof(double), _stream));
    RBF_CUDA(cudaMemsetAsync(_wcount.p, 0, _nOut * sizeof(int), _stream));
  }
  if (jit) { // the centre index is kept for the point queries of mapConsistentAt / mapConservativeAt (:285-287)
    double cMin[3], cMax[3];
    computeBounds(_centers.p, _nClusters, cMin, cMax, _stream);
    _centerGrid.build(_centers.p, _nClusters, cMin, cMax, 0.5 * _radius, _stream);
  }
  // sharded mapping, exchange mode: the weight sums of an output vertex span clusters of two ranks at a cut
  const auto weightsComplete = [&] {
    _group.allReduceSum(_wsum.p, _nOut, _stream);
    _group.allReduceSum(_wcount.p, _nOut, _stream);
  };
  if (_nClusters > 0) {
    pumBuildMembership(_centers.p, _nClusters, inGrid, outGrid, _radius, _mem, _scratch, _wsum.p, _wcount.p, _inPtr, _outPtr, _inRec, _outRec, _stream,
                       outSlab, exchangeMode() ? std::function<void()>(weightsComplete) : std::function<void()>(), _inverse);
  } else { // a rank of a sharded mapping without clusters still takes part in the collectives
    _mem = PumMembership();
    if (exchangeMode())
      weightsComplete();
  }
  trace.phase("membership");

  // partition-of-unity weight sums (PartitionOfUnityMapping.hpp:263-276, 290-341)
  auto throwUnassignedVertex = [&](int64_t bad) {
      double v[3];
      RBF_CUDA(cudaMemcpy(v, _outPtr + 3 * bad, sizeof(v), cudaMemcpyDeviceToHost));
      std::ostringstream os;
      os.precision(17);
      os << "Output vertex " << v[0] << " " << v[1];
      if (dim == 3)
        os << " " << v[2];
      os << " of mesh \"" << (cons ? "input" : "output")
         << "\" could not be assigned to any cluster in the rbf-pum mapping. This probably means that the meshes do not match well geometry-wise: "
            "Visualize the exported preCICE meshes to confirm. If the meshes are fine geometry-wise, you can try to increase the number of "
            "\"vertices-per-cluster\" (default is 50), the \"relative-overlap\" (default is 0.15), or disable the option \"project-to-input\". "
            "These options are only valid for the <mapping:rbf-pum-direct/> tag.";
      throw MapError(RBFMAP_ERR_UNASSIGNED, os.str());
  };
  if (_group.active()) {
    pumMarkFirstUnassigned(_wsum.p, _wcount.p, _nOut, _stream, _outPtr, outSlab, agree.p);
  } else {
    const int64_t bad = pumFirstUnassigned(_wsum.p, _wcount.p, _nOut, _stream, _outPtr);
    if (bad >= 0)
      throwUnassignedVertex(bad);
  }

  // deterministic mode: the weight sums (accumulated with atomics above) are recomputed per vertex in cluster order and
  // the normalised weights of the records rebuilt from them; the blend of map() uses the same vertex -> entries index
  if (_cfg.deterministic && _nClusters > 0 && !jit) {
    if (exchangeMode())
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "the deterministic blend is not available in the exchange shard mode (the all-reduce order is NCCL's).");
    PumVertexIndex outIndex;
    pumBuildVertexIndex(_mem.outIds.p, _mem.sumOut, _nOut, outIndex, _stream);
    PumSolveArgs a0 = solveArgs();
    pumDeterministicWeightSums(a0, _nClusters, outIndex, _wsum.p, _stream);
    pumBuildOutRecords(a0, _nClusters, _outRec.p, _stream);
    if (cons)
      pumBuildVertexIndex(_mem.inIds.p, _mem.sumIn, _nIn, _blendIndex, _stream); // the result lives on the in-vertices
    else
      _blendIndex = std::move(outIndex);
  }
  trace.phase("weights");
  // size buckets + statistics (device pass over the per-cluster counts)
  if (_nClusters > 0) {
    pumClassify(_mem, _nClusters, _buckets, stats, _stream);
  } else {
    for (int b = 0; b <= kNumBuckets; ++b)
      _buckets.start[b] = 0;
  }
  stats.n_clusters        = _nClusters;
  stats.n_clusters_global = _nClustersGlobal;
  stats.cluster_radius    = _radius;
  trace.phase("classify");
  // cluster-size limits of the factor / inverse kernels; on a sharded mapping the largest cluster of ALL ranks decides
  const auto checkClusterSize = [&](int64_t maxN) {
    if (_inverse && maxN > kPumMaxInverseVertices)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct with a basis function that is not strictly positive definite supports clusters of up to " +
                                                 std::to_string(kPumMaxInverseVertices) + " vertices (found " + std::to_string(maxN) +
                                                 "); reduce \"vertices-per-cluster\".");
    if (maxN > kPumMaxClusterVertices)
      throw MapError(RBFMAP_ERR_UNSUPPORTED, "rbf-pum-direct: a cluster with " + std::to_string(maxN) + " vertices exceeds the supported cluster size of " +
                                                 std::to_string(kPumMaxClusterVertices) + " vertices; reduce \"vertices-per-cluster\".");
  };
  const bool tooLarge = stats.max_n > (_inverse ? kPumMaxInverseVertices : kPumMaxClusterVertices);
  if (!_group.active())
    checkClusterSize(stats.max_n);
  // separated polynomial + assembly/factorisation of all local systems
  _poly.alloc(std::max<int64_t>(_nClusters, 1));
  if (!tooLarge && _mat.n < static_cast<size_t>(std::max<int64_t>(_mem.sumTri, 1))) {
    RBF_CUDA(cudaStreamSynchronize(_stream));
    _mat.reserve(static_cast<size_t>(_mem.sumTri) + static_cast<size_t>(_mem.sumTri) / 16 + 2);
  }
  const PumSolveArgs args = solveArgs();
  trace.phase("alloc factors");
  if (_nClusters > 0 && !tooLarge)
    pumPolySetup(args, _nClusters, _cfg.polynomial, _stream);
  trace.phase("polySetup");
  DevBuf<int> failBuf;
  int        *failP = nullptr; // lowest failing cluster index or INT_MAX
  if (_group.active()) {
    failP = agree.p + 1;
    const int negMax = -static_cast<int>(stats.max_n);
    RBF_CUDA(cudaMemcpyAsync(agree.p + 2, &negMax, sizeof(int), cudaMemcpyHostToDevice, _stream));
  } else {
    failBuf.alloc(1);
    const int noFail = INT_MAX;
    RBF_CUDA(cudaMemcpyAsync(failBuf.p, &noFail, sizeof(int), cudaMemcpyHostToDevice, _stream));
    failP = failBuf.p;
  }
  EventTimer tf(_stream);
  tf.start();
  if (tooLarge)
    ; // sharded mapping: this rank joins the agreement below and raises the size error with the others
  else if (_inverse)
    pumInverse(args, _buckets, failP, _stream);
  else
    pumFactor(args, _buckets, failP, _stream);
  stats.ms_assembly_factor = tf.stop();
  trace.phase("assembly + factorisation");
  int failed = INT_MAX;
  if (_group.active()) {
    _group.groupStart();
    _group.allReduceMin(agree.p, 3, _stream);
    _group.allReduceSum(agree.p + 3, 1, _stream);
    _group.groupEnd();
    int h[4];
    RBF_CUDA(cudaMemcpyAsync(h, agree.p, sizeof(h), cudaMemcpyDeviceToHost, _stream));
    RBF_CUDA(cudaStreamSynchronize(_stream));
    checkClusterSize(-h[2]);
    failed = h[1];
    if (_group.connected()) {
      _nClustersGlobal        = h[3];
      stats.n_clusters_global = _nClustersGlobal;
      if (_nClustersGlobal == 0)
        throw MapError(RBFMAP_ERR_EMPTY, "Too many vertices have been filtered out."); // CreateClustering.hpp:502, decided by all ranks together
    }
    // the reference builds the cluster solvers (singular matrix) before it assigns the weights (unassigned vertex)
    if (failed == INT_MAX && h[0] != 0x7f7f7f7f)
      throwUnassignedVertex(h[0]);
  } else {
    RBF_CUDA(cudaMemcpy(&failed, failBuf.p, sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (failed != INT_MAX) {
    // RadialBasisFctSolver.hpp:360-365
    const char *inName = cons ? "output" : "input", *outName = cons ? "input" : "output";
    throw MapError(RBFMAP_ERR_SINGULAR, std::string("The interpolation matrix of the RBF mapping from mesh \"") + inName + "\" to mesh \"" + outName +
                                            "\" is not invertable. This means that the mapping problem is not well-posed. "
                                            "Please check if your coupling meshes are correct (e.g. no vertices are duplicated) or reconfigure "
                                            "your basis-function (e.g. reduce the support-radius).");
  }
  // execution-mode minimal-compute: evaluate and store A_c once (BatchedRBFSolver.hpp:196-239, 324); map() streams them
  _evalReady = false;
  if (_cfg.execution_mode == RBFMAP_MINIMAL_COMPUTE && !jit && _nClusters > 0) {
    EventTimer te(_stream);
    te.start();
    const bool    lanesAreOut = !cons; // the side the map kernel assigns to its threads (pum_eval.cu)
    const int64_t total       = pumEvalOffsets(_mem, _nClusters, lanesAreOut, _evalOff, _stream);
    if (_eval.n < static_cast<size_t>(std::max<int64_t>(total, 1))) {
      RBF_CUDA(cudaStreamSynchronize(_stream));
      _eval.reserve(static_cast<size_t>(total) + static_cast<size_t>(total) / 16 + 2);
    }
    pumBuildEvalMatrices(args, _nClusters, lanesAreOut, reinterpret_cast<const long long *>(_evalOff.p), _eval.p, _stream);
    stats.ms_evaluation_offline = te.stop();
    stats.evaluation_bytes      = total * static_cast<int64_t>(sizeof(double));
    _evalReady                  = true;
    trace.phase("evaluation matrices");
  }
  stats.device_bytes       = stats.evaluation_bytes + static_cast<int64_t>(_inXyz.bytes() + _outXyz.bytes() + _centers.bytes() + _mem.bytes() + static_cast<size_t>(std::max<int64_t>(_mem.sumTri, 1)) * sizeof(double) + _poly.bytes() + _wsum.bytes() + _wcount.bytes() + _outRec.bytes() + _inRec.bytes());
  stats.ms_compute_kernels = whole.stop();
  stats.kernel_launches    = g_launches;
  _computed                = true;
}

void PumMapping::map(bool conservative, const double *dIn, int dims, double *dOut)
{
  if (!_computed)
    throw MapError(RBFMAP_ERR_STATE, "The mapping has not been computed; call computeMapping() first.");
  if (dims < 1)
    throw MapError(RBFMAP_ERR_INVALID, "data dimension has to be positive");
  const int        launchesBefore = g_launches;
  AllocStreamScope allocScope(_stream);
  EventTimer       t(_stream);
  t.start();
  // result lives on the solver-side outMesh (consistent) or inMesh (conservative)
  const int64_t nRes = conservative ? _nIn : _nOut;
  if (nRes > 0)
    RBF_CUDA(cudaMemsetAsync(dOut, 0, static_cast<size_t>(nRes) * dims * sizeof(double), _stream));
  if (_nClusters > 0) {
    PumSolveArgs   args = solveArgs();
    DevBuf<double> stage;
    const bool     det = _cfg.deterministic && _blendIndex.nVertices > 0;
    if (det) {
      stage.alloc(static_cast<size_t>(std::max<int64_t>(conservative ? _mem.sumIn : _mem.sumOut, 1)) * dims);
      args.blendStage = stage.p;
    }
    if (conservative)
      pumMapConservative(args, _buckets, dIn, dims, dOut, _stream);
    else
      pumMapConsistent(args, _buckets, dIn, dims, dOut, _stream);
    if (det) {
      pumBlendGather(stage.p, _blendIndex, dims, dOut, _stream);
      RBF_CUDA(cudaStreamSynchronize(_stream)); // `stage` is released on return
    }
  }
  stats.ms_collectives = 0.f;
  if (exchangeMode()) { // every cluster has one owner: the partial blends of the ranks add up to the result
    EventTimer tc(_stream);
    tc.start();
    _group.allReduceSum(dOut, nRes * dims, _stream);
    stats.ms_collectives = tc.stop();
  }
  stats.ms_map_kernels = t.stop();
  stats.kernel_launches += g_launches - launchesBefore;
}

// ---- just-in-time mapping (host pointers) ----
void PumMapping::throwUnassigned(const double *dCoords, int64_t bad, bool cons) const
{
  double v[3];
  RBF_CUDA(cudaMemcpy(v, dCoords + 3 * bad, sizeof(v), cudaMemcpyDeviceToHost));
  std::ostringstream os;
  os.precision(17);
  os << "Output vertex " << v[0] << " " << v[1];
  if (_cfg.dimensions == 3)
    os << " " << v[2];
  os << " of mesh \"" << (cons ? "input" : "output")
     << "\" could not be assigned to any cluster in the rbf-pum mapping. This probably means that the meshes do not match well geometry-wise: "
        "Visualize the exported preCICE meshes to confirm. If the meshes are fine geometry-wise, you can try to increase the number of "
        "\"vertices-per-cluster\" (default is 50), the \"relative-overlap\" (default is 0.15), or disable the option \"project-to-input\". "
        "These options are only valid for the <mapping:rbf-pum-direct/> tag.";
  throw MapError(RBFMAP_ERR_UNASSIGNED, os.str());
}

void PumMapping::jitUpdateCache(JitCache &c, const double *hValues)
{
  requireJit();
  checkCache(c);
  if (isConservative())
    throw MapError(RBFMAP_ERR_STATE, "updateMappingDataCache belongs to consistent (read) just-in-time mappings.");
  AllocStreamScope allocScope(_stream);
  DevBuf<double>   dIn;
  dIn.alloc(static_cast<size_t>(std::max<int64_t>(_nIn, 1)) * c.dims);
  if (_nIn > 0)
    RBF_CUDA(cudaMemcpyAsync(dIn.p, hValues, static_cast<size_t>(_nIn) * c.dims * sizeof(double), cudaMemcpyHostToDevice, _stream));
  if (_nClusters > 0)
    pumJitUpdateCache(solveArgs(), _buckets, dIn.p, c.dims, c.coef.p, c.poly.p, _stream);
  RBF_CUDA(cudaStreamSynchronize(_stream));
}

void PumMapping::jitMapConsistentAt(JitCache &c, const double *hCoords, int64_t nPts, double *hOut)
{
  requireJit();
  checkCache(c);
  if (nPts <= 0)
    return;
  AllocStreamScope allocScope(_stream);
  DevBuf<double>   dX, dOut;
  DevBuf<int>      first;
  dX.alloc(3 * nPts);
  dOut.alloc(static_cast<size_t>(nPts) * c.dims);
  first.alloc(1);
  RBF_CUDA(cudaMemcpyAsync(dX.p, hCoords, 3 * nPts * sizeof(double), cudaMemcpyHostToDevice, _stream));
  RBF_CUDA(cudaMemsetAsync(first.p, 0x7f, sizeof(int), _stream));
  RBF_CUDA(cudaMemsetAsync(dOut.p, 0, static_cast<size_t>(nPts) * c.dims * sizeof(double), _stream)); // values.setZero() :458
  bool clustered = false;
  if (nPts >= 4096 && stats.max_n <= 128) {
    // large point sets: bin the points, list them per cluster and evaluate cluster by cluster like the ordinary map
    double pMin[3], pMax[3];
    computeBounds(dX.p, nPts, pMin, pMax, _stream);
    BinGrid ptGrid;
    ptGrid.build(dX.p, nPts, pMin, pMax, 0.5 * _radius, _stream);
    DevBuf<double>  wsum;
    DevBuf<int>     wcount, nP, ptIds;
    DevBuf<int64_t> ptOff;
    wsum.alloc(nPts);
    wcount.alloc(nPts);
    RBF_CUDA(cudaMemsetAsync(wsum.p, 0, nPts * sizeof(double), _stream));
    RBF_CUDA(cudaMemsetAsync(wcount.p, 0, nPts * sizeof(int), _stream));
    clustered = pumBuildPointMembership(_centers.p, _nClusters, ptGrid, _radius, _scratch, nP, ptOff, ptIds, wsum.p, wcount.p, _stream);
    if (clustered) {
      const int64_t bad = pumFirstUnassigned(wsum.p, wcount.p, nPts, _stream);
      if (bad >= 0)
        throwUnassigned(dX.p, bad, false);
      pumJitEvaluateClustered(solveArgs(), _nClusters, static_cast<int>(stats.max_n), dX.p, nP.p, reinterpret_cast<const long long *>(ptOff.p), ptIds.p, wsum.p,
                              wcount.p, c.dims, c.coef.p, c.poly.p, dOut.p, _stream);
      RBF_CUDA(cudaStreamSynchronize(_stream)); // the temporaries of this scope are in use until here
    }
  }
  if (!clustered)
    pumJitInterpolateAt(solveArgs(), _centerGrid.view(), dX.p, nPts, c.dims, c.coef.p, c.poly.p, dOut.p, first.p, _stream);
  int f = 0;
  RBF_CUDA(cudaMemcpyAsync(&f, first.p, sizeof(int), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaMemcpyAsync(hOut, dOut.p, static_cast<size_t>(nPts) * c.dims * sizeof(double), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  if (f != 0x7f7f7f7f)
    throwUnassigned(dX.p, f, false);
}

void PumMapping::jitMapConservativeAt(JitCache &c, const double *hCoords, int64_t nPts, const double *hSource)
{
  requireJit();
  checkCache(c);
  if (!isConservative())
    throw MapError(RBFMAP_ERR_STATE, "mapConservativeAt belongs to conservative (write) just-in-time mappings.");
  if (nPts <= 0)
    return;
  AllocStreamScope allocScope(_stream);
  DevBuf<double>   dX, dU;
  DevBuf<int>      first;
  dX.alloc(3 * nPts);
  dU.alloc(static_cast<size_t>(nPts) * c.dims);
  first.alloc(1);
  RBF_CUDA(cudaMemcpyAsync(dX.p, hCoords, 3 * nPts * sizeof(double), cudaMemcpyHostToDevice, _stream));
  RBF_CUDA(cudaMemcpyAsync(dU.p, hSource, static_cast<size_t>(nPts) * c.dims * sizeof(double), cudaMemcpyHostToDevice, _stream));
  RBF_CUDA(cudaMemsetAsync(first.p, 0x7f, sizeof(int), _stream));
  pumJitAccumulateAt(solveArgs(), _centerGrid.view(), dX.p, nPts, c.dims, dU.p, c.coef.p, c.poly.p, first.p, _stream);
  int f = 0;
  RBF_CUDA(cudaMemcpyAsync(&f, first.p, sizeof(int), cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
  if (f != 0x7f7f7f7f)
    throwUnassigned(dX.p, f, true);
}

void PumMapping::jitComplete(JitCache &c, double *hOut)
{
  requireJit();
  checkCache(c);
  if (!isConservative())
    throw MapError(RBFMAP_ERR_STATE, "completeJustInTimeMapping belongs to conservative (write) just-in-time mappings.");
  if (_nIn == 0)
    return;
  AllocStreamScope allocScope(_stream);
  DevBuf<double>   dOut;
  const size_t     bytes = static_cast<size_t>(_nIn) * c.dims * sizeof(double);
  dOut.alloc(static_cast<size_t>(_nIn) * c.dims);
  RBF_CUDA(cudaMemcpyAsync(dOut.p, hOut, bytes, cudaMemcpyHostToDevice, _stream)); // the result is accumulated (:248-252)
  if (_nClusters > 0) {
    PumSolveArgs args = solveArgs();
    args.cacheAu      = c.coef.p;
    args.cachePoly    = c.poly.p;
    pumMapConservative(args, _buckets, nullptr, c.dims, dOut.p, _stream);
  }
  RBF_CUDA(cudaMemcpyAsync(hOut, dOut.p, bytes, cudaMemcpyDeviceToHost, _stream));
  RBF_CUDA(cudaStreamSynchronize(_stream));
}

} // namespace

std::unique_ptr<DeviceMapping> makePumMapping(const rbfmap_config &cfg) { return std::make_unique<PumMapping>(cfg); }

// ---- inspection helpers used by capi.cu ----
bool pumInspect(DeviceMapping *m, int64_t &nClusters, const double *&centers, const PumMembership *&mem)
{
  auto *p = dynamic_cast<PumMapping *>(m);
  if (!p)
    return false;
  nClusters = p->nClusters();
  centers   = p->centers().p;
  mem       = &p->membership();
  return true;
}

} // namespace rbfmap
