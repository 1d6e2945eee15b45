// rbfmap_mapping.hpp — header-only C++17 host side above the C ABI of rbfmap.h.
//
// The reference is C++: this is the mirror of its mapping interface for the RBF hot path, with the reference's names,
// argument meaning and error behaviour, so that code (and tests) written against `precice::mapping::Mapping`
// (src/mapping/Mapping.hpp:16-362) reads the same against librbfmap.so:
//
//   auto in  = std::make_shared<rbfmap::mesh::Mesh>("InMesh", 2);   in->createVertex({0.0, 0.0}); ...
//   rbfmap::mapping::PartitionOfUnityMapping mapping(rbfmap::mapping::Mapping::CONSISTENT, 2,
//                                                    rbfmap::mapping::CompactPolynomialC0(3.0),
//                                                    rbfmap::mapping::Polynomial::SEPARATE, 5, 0.4, false);
//   mapping.setMeshes(in, out);
//   mapping.computeMapping();
//   mapping.map(rbfmap::time::Sample{1, values}, result);
//
// It stands in for the pieces of preCICE the adapter of INTEGRATION.md touches (mesh::Mesh, time::Sample, the Mapping
// base class) where preCICE itself is not available; inside preCICE the adapter class of INTEGRATION.md §2 derives
// from the real precice::mapping::Mapping instead. No CUDA or torch types: it compiles with any C++17 compiler and links
// against librbfmap.so only. Citations are relative to /root/reference/src/.
#pragma once

#include "rbfmap.h"

#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rbfmap {

/// precice::Error (precice/Exceptions.hpp:9-14): what PRECICE_CHECK / PRECICE_ERROR throw; carries the C ABI status as well
class Error : public std::runtime_error {
public:
  Error(int status, const std::string &what) : std::runtime_error(what), status(status) {}
  int status; ///< rbfmap_status
};

namespace mesh {

/// The part of mesh::Mesh (mesh/Mesh.hpp) a mapping reads: vertex coordinates (Vertex::rawCoords(), z = 0 in 2-D), the
/// tags set by the re-partitioning hooks, the owner flags, and the connectivity the scaled-consistent constraints integrate
class Mesh {
public:
  Mesh(std::string name, int dimensions) : _name(std::move(name)), _dimensions(dimensions) {}

  /// mesh::Mesh::createVertex (mesh/Mesh.hpp:91): returns the vertex ID
  int createVertex(std::initializer_list<double> coords)
  {
    std::array<double, 3> c{0.0, 0.0, 0.0};
    int                   d = 0;
    for (double v : coords)
      if (d < 3)
        c[d++] = v;
    _coords.push_back(c);
    _tags.push_back(0);
    _owner.push_back(1);
    return static_cast<int>(_coords.size()) - 1;
  }
  int createVertex(const std::array<double, 3> &c) { return createVertex({c[0], c[1], c[2]}); }
  void createEdge(int a, int b) { _edges.insert(_edges.end(), {a, b}); }
  void createTriangle(int a, int b, int c) { _triangles.insert(_triangles.end(), {a, b, c}); }
  void createTetrahedron(int a, int b, int c, int d) { _tetrahedra.insert(_tetrahedra.end(), {a, b, c, d}); }

  const std::string &getName() const { return _name; }
  int                getDimensions() const { return _dimensions; }
  std::size_t        nVertices() const { return _coords.size(); }
  bool               empty() const { return _coords.empty(); }
  /// contiguous double[n][3]: the layout rbfmap_compute_mapping takes
  const double *rawCoords() const { return _coords.empty() ? nullptr : _coords.front().data(); }
  const std::array<double, 3> &vertexCoords(std::size_t i) const { return _coords[i]; }
  bool isTagged(std::size_t i) const { return _tags[i] != 0; }   ///< Vertex::isTagged
  void setOwner(std::size_t i, bool owner) { _owner[i] = owner; } ///< Vertex::setOwner
  std::vector<unsigned char>       &tags() { return _tags; }
  const std::vector<unsigned char> &owners() const { return _owner; }
  const std::vector<int32_t>       &edges() const { return _edges; }
  const std::vector<int32_t>       &triangles() const { return _triangles; }
  const std::vector<int32_t>       &tetrahedra() const { return _tetrahedra; }

private:
  std::string                        _name;
  int                                _dimensions;
  std::vector<std::array<double, 3>> _coords;
  std::vector<unsigned char>         _tags, _owner;
  std::vector<int32_t>               _edges, _triangles, _tetrahedra;
};
using PtrMesh = std::shared_ptr<Mesh>;

} // namespace mesh

namespace time {
/// time::Sample (time/Sample.hpp:52-55): values[i * dataDims + c]
struct Sample {
  int                 dataDims = 1;
  std::vector<double> values;
};
} // namespace time

namespace mapping {

/// mapping/config/MappingConfigurationTypes.hpp:11-15
enum class Polynomial { ON = RBFMAP_POLYNOMIAL_ON, OFF = RBFMAP_POLYNOMIAL_OFF, SEPARATE = RBFMAP_POLYNOMIAL_SEPARATE };

/// A basis function object as the reference constructs it (mapping/impl/BasisFunctions.hpp:87-607): kind + its one argument
struct BasisFunction {
  int    kind;
  double supportRadius  = 0.0; ///< compact polynomials, compact thin-plate splines
  double shapeParameter = 0.0; ///< gaussian, (inverse) multiquadrics
};
inline BasisFunction CompactPolynomialC0(double supportRadius) { return {RBFMAP_COMPACT_POLYNOMIAL_C0, supportRadius, 0.0}; }
inline BasisFunction CompactPolynomialC2(double supportRadius) { return {RBFMAP_COMPACT_POLYNOMIAL_C2, supportRadius, 0.0}; }
inline BasisFunction CompactPolynomialC4(double supportRadius) { return {RBFMAP_COMPACT_POLYNOMIAL_C4, supportRadius, 0.0}; }
inline BasisFunction CompactPolynomialC6(double supportRadius) { return {RBFMAP_COMPACT_POLYNOMIAL_C6, supportRadius, 0.0}; }
inline BasisFunction CompactPolynomialC8(double supportRadius) { return {RBFMAP_COMPACT_POLYNOMIAL_C8, supportRadius, 0.0}; }
inline BasisFunction CompactThinPlateSplinesC2(double supportRadius) { return {RBFMAP_COMPACT_TPS_C2, supportRadius, 0.0}; }
inline BasisFunction ThinPlateSplines() { return {RBFMAP_THIN_PLATE_SPLINES, 0.0, 0.0}; }
inline BasisFunction VolumeSplines() { return {RBFMAP_VOLUME_SPLINES, 0.0, 0.0}; }
inline BasisFunction Multiquadrics(double shape) { return {RBFMAP_MULTIQUADRICS, 0.0, shape}; }
inline BasisFunction InverseMultiquadrics(double shape) { return {RBFMAP_INVERSE_MULTIQUADRICS, 0.0, shape}; }
inline BasisFunction Gaussian(double shape, double supportRadius = 0.0) { return {RBFMAP_GAUSSIAN, supportRadius, shape}; }

class Mapping;

/// impl::MappingDataCache (mapping/impl/MappingDataCache.hpp:23-71): the per-data state of a just-in-time mapping, created by
/// Mapping::initializeMappingDataCache and owned by the caller (precice/impl/ReadDataContext.cpp, WriteDataContext.cpp)
class MappingDataCache {
public:
  explicit MappingDataCache(int dataDim) : _dataDim(dataDim) {}
  MappingDataCache(const MappingDataCache &)            = delete;
  MappingDataCache &operator=(const MappingDataCache &) = delete;
  ~MappingDataCache() { rbfmap_cache_destroy(_c); }
  int getDataDimensions() const { return _dataDim; }

private:
  friend class Mapping;
  int           _dataDim;
  rbfmap_cache *_c = nullptr;
};

/// The interface of precice::mapping::Mapping (mapping/Mapping.hpp) for the three RBF mappings, on one B200
class Mapping {
public:
  /// mapping/Mapping.hpp:30-35 (same enumerator values as rbfmap_constraint)
  enum Constraint { CONSISTENT = 0, CONSERVATIVE = 1, SCALED_CONSISTENT_SURFACE = 2, SCALED_CONSISTENT_VOLUME = 3 };

  Mapping(const Mapping &)            = delete;
  Mapping &operator=(const Mapping &) = delete;
  Mapping &operator=(Mapping &&)      = delete; // mapping/Mapping.hpp:72
  virtual ~Mapping() { rbfmap_destroy(_h); }

  /// mapping/Mapping.cpp:28-34
  void setMeshes(const mesh::PtrMesh &input, const mesh::PtrMesh &output)
  {
    _input  = input;
    _output = output;
  }
  const mesh::PtrMesh &getInputMesh() const { return _input; }
  const mesh::PtrMesh &getOutputMesh() const { return _output; }
  Constraint           getConstraint() const { return _constraint; }
  int                  getDimensions() const { return _dimensions; }
  bool hasConstraint(Constraint c) const { return _constraint == c; }
  bool isScaledConsistent() const { return _constraint == SCALED_CONSISTENT_SURFACE || _constraint == SCALED_CONSISTENT_VOLUME; }
  bool hasComputedMapping() const { return rbfmap_has_computed_mapping(_h) != 0; }
  /// InitialGuessRequirement::Required for the iterative variant (PetRadialBasisFctMapping.hpp:176)
  bool requiresInitialGuess() const { return rbfmap_initial_guess_size(_h) > 0 || _cfg.variant == RBFMAP_GLOBAL_ITERATIVE; }
  std::string getName() const { return rbfmap_get_name(_h); }

  /// Mapping::computeMapping (mapping/Mapping.hpp:101)
  void computeMapping()
  {
    requireMeshes();
    if (isScaledConsistent()) { // Mapping.cpp:190-208: edges (2-D surface), triangles (2-D volume / 3-D surface), tetrahedra
      const bool volume = _constraint == SCALED_CONSISTENT_VOLUME;
      for (int which = 0; which < 2; ++which) {
        const mesh::Mesh           &m = which == 0 ? *_input : *_output;
        const int                   k = _dimensions == 2 ? (volume ? 3 : 2) : (volume ? 4 : 3);
        const std::vector<int32_t> &el = k == 2 ? m.edges() : (k == 3 ? m.triangles() : m.tetrahedra());
        check(rbfmap_set_connectivity(_h, which, el.data(), static_cast<int64_t>(el.size()) / k, k));
      }
    }
    check(rbfmap_compute_mapping(_h, _input->rawCoords(), static_cast<int64_t>(_input->nVertices()), _output->rawCoords(),
                                 static_cast<int64_t>(_output->nVertices())));
  }
  /// Mapping::clear (mapping/Mapping.hpp:134)
  void clear() { check(rbfmap_clear(_h)); }

  /// Mapping::map(const time::Sample &, Eigen::VectorXd &) (mapping/Mapping.cpp:158-173): dispatches on the constraint,
  /// including the scaled-consistent post-scaling; `output` is resized to dataDims x output vertices
  void map(const time::Sample &input, std::vector<double> &output)
  {
    requireMeshes();
    checkSample(input);
    output.assign(static_cast<std::size_t>(input.dataDims) * _output->nVertices(), 0.0);
    check(rbfmap_map(_h, input.values.data(), input.dataDims, output.data()));
  }
  /// Mapping::map(input, output, initialGuess) (mapping/Mapping.cpp:150-156): the caller keeps one guess per data
  void map(const time::Sample &input, std::vector<double> &output, std::vector<double> &initialGuess)
  {
    requireMeshes();
    checkSample(input);
    const std::size_t need = static_cast<std::size_t>(rbfmap_initial_guess_size(_h)) * input.dataDims;
    if (initialGuess.size() != need)
      initialGuess.assign(need, 0.0); // "hasInitialGuess() == false": start from zero
    output.assign(static_cast<std::size_t>(input.dataDims) * _output->nVertices(), 0.0);
    check(rbfmap_map_with_initial_guess(_h, input.values.data(), input.dataDims, output.data(), initialGuess.data()));
  }

  /// Mapping::tagMeshFirstRound / tagMeshSecondRound (mapping/Mapping.hpp:179-182, caller partition/ReceivedPartition.cpp:921-954):
  /// tags the vertices of the received mesh - input() for consistent, output() for conservative constraints - to keep
  void tagMeshFirstRound()
  {
    requireMeshes();
    mesh::Mesh &remote = hasConstraint(CONSERVATIVE) ? *_output : *_input;
    if (remote.empty())
      return;
    check(rbfmap_tag_mesh_first_round(_h, _input->rawCoords(), static_cast<int64_t>(_input->nVertices()), _output->rawCoords(),
                                      static_cast<int64_t>(_output->nVertices()), remote.tags().data()));
  }
  void tagMeshSecondRound()
  {
    requireMeshes();
    mesh::Mesh &remote = hasConstraint(CONSERVATIVE) ? *_output : *_input;
    if (remote.empty())
      return;
    check(rbfmap_tag_mesh_second_round(_h, remote.rawCoords(), static_cast<int64_t>(remote.nVertices()), remote.owners().data(), remote.tags().data()));
  }

  // ---- just-in-time mapping (mapping/Mapping.hpp:215-289; rbf-pum-direct only, PartitionOfUnityMapping.hpp:382-476) ----
  /// computeMapping() when the other mesh is a just-in-time mesh (mesh::MeshConfiguration::getJustInTimeMappingMesh): only
  /// input() (consistent / read) resp. output() (conservative / write) has to be set
  void computeMappingJustInTime()
  {
    const mesh::PtrMesh &m = hasConstraint(CONSERVATIVE) ? _output : _input;
    if (!m)
      throw Error(RBFMAP_ERR_STATE, "setMeshes() has to be called before the mapping is used.");
    check(rbfmap_compute_mapping_just_in_time(_h, m->rawCoords(), static_cast<int64_t>(m->nVertices())));
  }
  /// Mapping::initializeMappingDataCache (PartitionOfUnityMapping.hpp:424-434); also resets an initialised cache
  void initializeMappingDataCache(MappingDataCache &cache)
  {
    if (cache._c)
      check(rbfmap_cache_reset(_h, cache._c));
    else
      check(rbfmap_cache_create(_h, cache._dataDim, &cache._c));
  }
  /// Mapping::updateMappingDataCache (PartitionOfUnityMapping.hpp:436-448): values of input(), dataDim per vertex
  void updateMappingDataCache(MappingDataCache &cache, const std::vector<double> &in) { check(rbfmap_cache_update(_h, cache._c, in.data())); }
  /// Mapping::mapConsistentAt (PartitionOfUnityMapping.hpp:450-476): coordinates = nPoints x 3 (z = 0 in 2-D),
  /// values = nPoints x dataDim
  void mapConsistentAt(const std::vector<std::array<double, 3>> &coordinates, MappingDataCache &cache, std::vector<double> &values)
  {
    values.assign(coordinates.size() * cache._dataDim, 0.0);
    check(rbfmap_map_consistent_at(_h, cache._c, coordinates.empty() ? nullptr : coordinates.front().data(), static_cast<int64_t>(coordinates.size()), values.data()));
  }
  /// Mapping::mapConservativeAt (PartitionOfUnityMapping.hpp:382-409): accumulates `source` (nPoints x dataDim) into the cache
  void mapConservativeAt(const std::vector<std::array<double, 3>> &coordinates, const std::vector<double> &source, MappingDataCache &cache)
  {
    check(rbfmap_map_conservative_at(_h, cache._c, coordinates.empty() ? nullptr : coordinates.front().data(), static_cast<int64_t>(coordinates.size()), source.data()));
  }
  /// Mapping::completeJustInTimeMapping (PartitionOfUnityMapping.hpp:411-422): adds the mapped data to `result` (dataDim x output vertices)
  void completeJustInTimeMapping(MappingDataCache &cache, std::vector<double> &result) { check(rbfmap_complete_just_in_time_mapping(_h, cache._c, result.data())); }

  rbfmap_stats stats() const
  {
    rbfmap_stats st;
    rbfmap_get_stats(_h, &st);
    return st;
  }
  rbfmap_handle *handle() const { return _h; }

protected:
  Mapping(Constraint constraint, int dimensions, const rbfmap_config &cfg) : _constraint(constraint), _dimensions(dimensions), _cfg(cfg)
  {
    _cfg.constraint = static_cast<int>(constraint);
    _cfg.dimensions = dimensions;
    const int rc    = rbfmap_create(&_cfg, &_h);
    if (rc != RBFMAP_OK)
      throw Error(rc, rbfmap_last_error(nullptr)); // the PRECICE_CHECKs of the reference constructors
  }
  void check(int rc) const
  {
    if (rc != RBFMAP_OK)
      throw Error(rc, rbfmap_last_error(_h));
  }

private:
  void requireMeshes() const
  {
    if (!_input || !_output)
      throw Error(RBFMAP_ERR_STATE, "setMeshes() has to be called before the mapping is used.");
  }
  void checkSample(const time::Sample &s) const
  {
    if (s.dataDims < 1 || s.values.size() != static_cast<std::size_t>(s.dataDims) * _input->nVertices())
      throw Error(RBFMAP_ERR_INVALID, "The sample does not hold dataDims values per vertex of the input mesh.");
  }

  Constraint     _constraint;
  int            _dimensions;
  rbfmap_config  _cfg;
  rbfmap_handle *_h = nullptr;
  mesh::PtrMesh  _input, _output;
};

namespace detail {
inline rbfmap_config baseConfig(int variant, const BasisFunction &fn, Polynomial polynomial, int deviceId)
{
  rbfmap_config c;
  rbfmap_default_config(&c);
  c.variant         = variant;
  c.basis           = fn.kind;
  c.support_radius  = fn.supportRadius;
  c.shape_parameter = fn.shapeParameter;
  c.polynomial      = static_cast<int>(polynomial);
  c.device_id       = deviceId;
  return c;
}
} // namespace detail

/// <mapping:rbf-pum-direct> (mapping/PartitionOfUnityMapping.hpp:48-57; constructor arguments in the reference's order,
/// `computeEvaluationOffline` = <executor:cuda execution-mode="minimal-compute">)
class PartitionOfUnityMapping : public Mapping {
public:
  PartitionOfUnityMapping(Constraint constraint, int dimension, const BasisFunction &function, Polynomial polynomial, unsigned verticesPerCluster,
                          double relativeOverlap, bool projectToInput, bool computeEvaluationOffline = false, int gpuDeviceId = -1)
      : Mapping(constraint, dimension, config(function, polynomial, verticesPerCluster, relativeOverlap, projectToInput, computeEvaluationOffline, gpuDeviceId))
  {
  }

private:
  static rbfmap_config config(const BasisFunction &fn, Polynomial polynomial, unsigned vpc, double overlap, bool project, bool offline, int deviceId)
  {
    rbfmap_config c        = detail::baseConfig(RBFMAP_PUM_DIRECT, fn, polynomial, deviceId);
    c.vertices_per_cluster = vpc;
    c.relative_overlap     = overlap;
    c.project_to_input     = project ? 1 : 0;
    c.execution_mode       = offline ? RBFMAP_MINIMAL_COMPUTE : RBFMAP_MINIMAL_MEMORY;
    return c;
  }
};

/// <mapping:rbf-global-direct> (mapping/RadialBasisFctMapping.hpp:40-46)
class RadialBasisFctMapping : public Mapping {
public:
  RadialBasisFctMapping(Constraint constraint, int dimensions, const BasisFunction &function, std::array<bool, 3> deadAxis, Polynomial polynomial,
                        int gpuDeviceId = -1)
      : Mapping(constraint, dimensions, config(RBFMAP_GLOBAL_DIRECT, function, deadAxis, polynomial, gpuDeviceId))
  {
  }

protected:
  RadialBasisFctMapping(Constraint constraint, int dimensions, const rbfmap_config &cfg) : Mapping(constraint, dimensions, cfg) {}
  static rbfmap_config config(int variant, const BasisFunction &fn, std::array<bool, 3> deadAxis, Polynomial polynomial, int deviceId)
  {
    rbfmap_config c = detail::baseConfig(variant, fn, polynomial, deviceId);
    for (int d = 0; d < 3; ++d)
      c.dead_axis[d] = deadAxis[d] ? 1 : 0;
    return c;
  }
};

/// <mapping:rbf-global-iterative> (semantics of mapping/PetRadialBasisFctMapping.hpp:44-52: solver-rtol, polynomial,
/// rescaling by the interpolant of one, caller-owned initial guess)
class IterativeRadialBasisFctMapping : public RadialBasisFctMapping {
public:
  IterativeRadialBasisFctMapping(Constraint constraint, int dimensions, const BasisFunction &function, std::array<bool, 3> deadAxis, double solverRtol = 1e-9,
                                 Polynomial polynomial = Polynomial::SEPARATE, int gpuDeviceId = -1)
      : RadialBasisFctMapping(constraint, dimensions, iterativeConfig(function, deadAxis, solverRtol, polynomial, gpuDeviceId))
  {
  }

private:
  static rbfmap_config iterativeConfig(const BasisFunction &fn, std::array<bool, 3> deadAxis, double rtol, Polynomial polynomial, int deviceId)
  {
    rbfmap_config c = config(RBFMAP_GLOBAL_ITERATIVE, fn, deadAxis, polynomial, deviceId);
    c.solver_rtol   = rtol;
    return c;
  }
};

} // namespace mapping
} // namespace rbfmap
