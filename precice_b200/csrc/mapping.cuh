// mapping.cuh — host-side mapping objects behind the C ABI (one per XML tag of the hot path).
#pragma once

#include "basis.cuh"
#include "common.cuh"
#include "nccl_dyn.cuh"

#include <memory>
#include <string>

namespace rbfmap {

// Common interface; mirrors the virtuals of mapping::Mapping that the hot path implements
// (mapping/Mapping.hpp:101 computeMapping, :134 clear, :324/:337 mapConservative/mapConsistent, :197 getName).
class DeviceMapping {
public:
  explicit DeviceMapping(const rbfmap_config &cfg);
  virtual ~DeviceMapping();

  // input/output = Mapping::input()/output() coordinates on the device (AoS n x 3)
  virtual void        computeMapping(const double *dInputXyz, int64_t nInput, const double *dOutputXyz, int64_t nOutput) = 0;
  // host-pointer variant behind rbfmap_compute_mapping; returns false if the variant has no own implementation (the
  // caller then stages both meshes on the device and calls computeMapping). rbf-pum-direct overrides it to overlap the
  // upload of the second mesh with the work that only needs the first one.
  virtual bool computeMappingHost(const double * /*hInputXyz*/, int64_t /*nInput*/, const double * /*hOutputXyz*/, int64_t /*nOutput*/) { return false; }
  // ---- just-in-time mapping (Mapping.hpp: initializeMappingDataCache / updateMappingDataCache / mapConsistentAt /
  // mapConservativeAt / completeJustInTimeMapping); host pointers; only rbf-pum-direct implements it ----
  struct JitCache { // impl/MappingDataCache.hpp:23-71 on the device
    int                dims = 0;
    ScratchBuf<double> coef, poly;
  };
  [[noreturn]] static void noJit() { throw MapError(RBFMAP_ERR_UNSUPPORTED, "just-in-time mapping is implemented for rbf-pum-direct only"); }
  virtual void                      jitComputeMapping(const double * /*hMeshXyz*/, int64_t /*n*/) { noJit(); }
  virtual std::unique_ptr<JitCache> jitCreateCache(int /*dims*/) { noJit(); }
  virtual void                      jitResetCache(JitCache &) { noJit(); }
  virtual void                      jitUpdateCache(JitCache &, const double * /*hValues*/) { noJit(); }
  virtual void jitMapConsistentAt(JitCache &, const double * /*hCoords*/, int64_t /*nPts*/, double * /*hOut*/) { noJit(); }
  virtual void jitMapConservativeAt(JitCache &, const double * /*hCoords*/, int64_t /*nPts*/, const double * /*hSource*/) { noJit(); }
  virtual void jitComplete(JitCache &, double * /*hOut, accumulated*/) { noJit(); }
  virtual void        mapConsistent(const double *dIn, int dims, double *dOut)                                         = 0;
  virtual void        mapConservative(const double *dIn, int dims, double *dOut)                                       = 0;
  virtual void        clear()                                                                                          = 0;
  virtual std::string name() const                                                                                     = 0;
  // joins a group of `world` processes (one GPU each) that execute the same calls on replicated inputs and shard the work:
  // rbf-global-iterative shards the rows of its CG operator (global_iterative.cu), rbf-pum-direct its clusters (pum.cu)
  virtual void distributedInit(const char * /*ncclUniqueId, 128 bytes*/, int /*rank*/, int /*world*/)
  {
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "this mapping variant does not shard across devices: rbf-global-direct stays on one device");
  }
  const ShardGroup &group() const { return _group; }
  // Host -> device copy of `count` doubles that every rank of the group holds identically: each rank uploads only its
  // 1/world slice over PCIe and the slices are all-gathered over NVLink. `dDst` must hold group().padded(count) doubles.
  // Enqueued on `s`; the host buffer may be reused after `s` has been synchronised.
  void uploadReplicated(double *dDst, const double *hSrc, int64_t count, cudaStream_t s) const;
  // true if map() leaves on this rank only a part of the result and the complete result is the sum over the ranks of the
  // group (rbf-pum-direct, duplicate shard mode); false if every rank holds the complete result
  virtual bool mapResultIsPartial() const { return false; }

  // ---- re-partitioning hooks (Mapping::tagMeshFirstRound / tagMeshSecondRound, Mapping.hpp:179-182), host pointers ----
  // Marks (tags[i] = 1) the vertices of the REMOTE mesh - Mapping::input() for consistent, Mapping::output() for
  // conservative constraints - that a rank holding the LOCAL (other) mesh needs; untouched entries keep their value.
  // Default = RadialBasisFctBaseMapping.hpp:124-190 (the two global variants); rbf-pum-direct overrides the first round
  // (PartitionOfUnityMapping.hpp:479-533) and does nothing in the second (:535-539).
  virtual void tagMeshFirstRound(const double *hFilterXyz, int64_t nFilter, const double *hOtherXyz, int64_t nOther, unsigned char *hTags);
  virtual void tagMeshSecondRound(const double *hFilterXyz, int64_t nFilter, const unsigned char *hOwner, unsigned char *hTags);
  // impl::estimateClusterRadius (impl/CreateClustering.hpp:223-278) on an arbitrary mesh and bounding box; PUM only
  virtual double estimateClusterRadiusHost(const double * /*hXyz*/, int64_t /*n*/, const double * /*bbMin*/, const double * /*bbMax*/)
  {
    throw MapError(RBFMAP_ERR_UNSUPPORTED, "the cluster radius belongs to rbf-pum-direct");
  }

  // ---- caller-owned initial guess of iterative mappings (Mapping::requiresInitialGuess / initialGuess, Mapping.hpp) ----
  virtual int64_t initialGuessSize() const { return 0; }                      // per data dimension; 0 = none needed
  virtual void    loadInitialGuess(const double * /*hGuess*/, int /*dims*/) {} // before map()
  virtual void    storeInitialGuess(double * /*hGuess*/, int /*dims*/) {}      // after map()

  bool                 hasComputedMapping() const { return _computed; }
  bool                 isConservative() const { return _cfg.constraint == RBFMAP_CONSERVATIVE; }
  const rbfmap_config &config() const { return _cfg; }
  cudaStream_t         stream() const { return _stream; }
  int                  device() const { return _device; }
  int64_t              nInput() const { return _nInput; }
  int64_t              nOutput() const { return _nOutput; }
  rbfmap_stats         stats;

protected:
  // tags the vertices inside [lo - margin, hi + margin] (inclusive, BoundingBox::contains / bgi::intersects), first `dim` axes
  void tagInsideBox(const double *dXyz, int64_t n, const double lo[3], const double hi[3], double margin, unsigned char *hTags);
  rbfmap_config _cfg;
  RbfParams     _rbf;
  ShardGroup    _group;
  cudaStream_t  _stream   = nullptr;
  int           _device   = 0; // resolved CUDA device ordinal of this mapping (gpu-device-id)
  bool          _computed = false;
  int64_t       _nInput = 0, _nOutput = 0;
};

std::unique_ptr<DeviceMapping> makePumMapping(const rbfmap_config &cfg);
std::unique_ptr<DeviceMapping> makeGlobalDirectMapping(const rbfmap_config &cfg);
std::unique_ptr<DeviceMapping> makeGlobalIterativeMapping(const rbfmap_config &cfg);

// Phase tracing for development: RBFMAP_TRACE=1 prints the wall time of every phase of computeMapping/map to stderr
// (a stream synchronisation is inserted at each phase boundary, so traced runs are slower than untraced ones).
struct PhaseTrace {
  bool         on;
  cudaStream_t s;
  double       t0;
  const char  *what;
  static double now();
  PhaseTrace(const char *what_, cudaStream_t stream);
  void phase(const char *name);
};

// simple CUDA-event stopwatch on a stream
struct EventTimer {
  cudaEvent_t  a = nullptr, b = nullptr;
  cudaStream_t s;
  explicit EventTimer(cudaStream_t stream)
      : s(stream)
  {
    RBF_CUDA(cudaEventCreate(&a));
    RBF_CUDA(cudaEventCreate(&b));
  }
  ~EventTimer()
  {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
  void  start() { RBF_CUDA(cudaEventRecord(a, s)); }
  float stop()
  {
    RBF_CUDA(cudaEventRecord(b, s));
    RBF_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    RBF_CUDA(cudaEventElapsedTime(&ms, a, b));
    return ms;
  }
};

} // namespace rbfmap
