/* rbfmap.h — C ABI of the B200-native radial-basis-function data-mapping backend.
 *
 * This is the drop-in boundary for preCICE's RBF mapping hot path: everything a
 * `precice::mapping::Mapping` backend for rbf-pum-direct / rbf-global-direct /
 * rbf-global-iterative needs, as plain pointers and sizes (no C++/torch types), so that the
 * reference can bind it from `src/mapping` with a thin adapter class.
 * All citations are relative to /root/reference/src/ (preCICE 3.4.1).
 *
 * Conventions
 *  - coordinates: `double xyz[n][3]`, exactly `mesh::Vertex::rawCoords()` (mesh/Vertex.hpp:75-77);
 *    2-D meshes carry z = 0.
 *  - values: vertex-major interleaved `values[i*dims + c]` as in `time::Sample`
 *    (time/Sample.hpp:52-55, mapping/RadialBasisFctMapping.hpp:257-261).
 *  - `input` / `output` mean Mapping::input() / Mapping::output(); the swap for conservative
 *    constraints happens inside (mapping/PartitionOfUnityMapping.hpp:190-198).
 *  - every call is synchronous and returns RBFMAP_OK or an error code; the message (same texts
 *    as the reference's PRECICE_CHECKs) is available from rbfmap_last_error().
 *  - there is no CPU fallback: without a CUDA device every compute call fails with
 *    RBFMAP_ERR_CUDA.
 */
#ifndef RBFMAP_H
#define RBFMAP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mapping/config/MappingConfigurationTypes.hpp:11-15 */
enum rbfmap_polynomial { RBFMAP_POLYNOMIAL_ON = 0, RBFMAP_POLYNOMIAL_OFF = 1, RBFMAP_POLYNOMIAL_SEPARATE = 2 };

/* mapping/config/MappingConfigurationTypes.hpp:17-29 (functors: mapping/impl/BasisFunctions.hpp:87-607) */
enum rbfmap_basis {
  RBFMAP_COMPACT_POLYNOMIAL_C0 = 0,
  RBFMAP_COMPACT_POLYNOMIAL_C2 = 1,
  RBFMAP_COMPACT_POLYNOMIAL_C4 = 2,
  RBFMAP_COMPACT_POLYNOMIAL_C6 = 3,
  RBFMAP_COMPACT_POLYNOMIAL_C8 = 4,
  RBFMAP_THIN_PLATE_SPLINES    = 5,
  RBFMAP_MULTIQUADRICS         = 6,
  RBFMAP_INVERSE_MULTIQUADRICS = 7,
  RBFMAP_VOLUME_SPLINES        = 8,
  RBFMAP_GAUSSIAN              = 9,
  RBFMAP_COMPACT_TPS_C2        = 10
};

/* mapping/Mapping.hpp:30-35 */
enum rbfmap_constraint {
  RBFMAP_CONSISTENT                = 0,
  RBFMAP_CONSERVATIVE              = 1,
  RBFMAP_SCALED_CONSISTENT_SURFACE = 2,
  RBFMAP_SCALED_CONSISTENT_VOLUME  = 3
};

/* the three XML tags of the hot path (mapping/config/MappingConfiguration.cpp:834-914) */
enum rbfmap_variant { RBFMAP_GLOBAL_DIRECT = 0, RBFMAP_GLOBAL_ITERATIVE = 1, RBFMAP_PUM_DIRECT = 2 };

/* <executor:cuda execution-mode="minimal-memory|minimal-compute"> (mapping/config/MappingConfiguration.cpp:328-331,
 * 555-559) = constructor argument computeEvaluationOffline of PartitionOfUnityMapping (PartitionOfUnityMapping.hpp:48-57,
 * 145; used at BatchedRBFSolver.hpp:196, 239, 324). minimal-memory re-evaluates the evaluation matrix in every map();
 * minimal-compute stores the per-cluster evaluation operator at computeMapping() and map() becomes an HBM stream. */
enum rbfmap_execution_mode { RBFMAP_MINIMAL_MEMORY = 0, RBFMAP_MINIMAL_COMPUTE = 1 };

/* How rbf-pum-direct shards ONE mapping over the ranks joined by rbfmap_distributed_init:
 * DUPLICATE: every rank owns a slab of output vertices and processes every cluster that touches one of them (boundary
 *            clusters are factorised on both sides of a cut); no collective in map(); consistent constraints only.
 * EXCHANGE : every cluster belongs to exactly one rank; partial sums of the blend (and, at computeMapping, the partition-
 *            of-unity weight sums) are all-reduced with NCCL over NVLink; every rank returns the complete result. */
enum rbfmap_shard_mode { RBFMAP_SHARD_DUPLICATE = 0, RBFMAP_SHARD_EXCHANGE = 1 };

enum rbfmap_status {
  RBFMAP_OK              = 0,
  RBFMAP_ERR_INVALID     = 1, /* bad argument / configuration (PRECICE_CHECK at construction) */
  RBFMAP_ERR_CUDA        = 2, /* no device, CUDA runtime failure */
  RBFMAP_ERR_SINGULAR    = 3, /* RadialBasisFctSolver.hpp:360-365 "... is not invertable ..." */
  RBFMAP_ERR_UNASSIGNED  = 4, /* PartitionOfUnityMapping.hpp:314-318 "Output vertex ... could not be assigned ..." */
  RBFMAP_ERR_EMPTY       = 5, /* impl/CreateClustering.hpp:502 "Too many vertices have been filtered out." */
  RBFMAP_ERR_STATE       = 6, /* map before computeMapping, recompute without clear (PartitionOfUnityMapping.hpp:188) */
  RBFMAP_ERR_UNSUPPORTED = 7,
  RBFMAP_ERR_NOT_CONVERGED = 8
};

/* Constructor arguments of the three mapping classes:
 *   RadialBasisFctMapping   (mapping/RadialBasisFctMapping.hpp:40-46)
 *   PartitionOfUnityMapping (mapping/PartitionOfUnityMapping.hpp:48-57)
 * plus the basis-function constructor argument (mapping/impl/BasisFunctions.hpp) and the
 * <executor:cuda gpu-device-id=...> attribute (mapping/config/MappingConfiguration.cpp:283-338). */
typedef struct rbfmap_config {
  int      variant;              /* rbfmap_variant */
  int      constraint;           /* rbfmap_constraint */
  int      dimensions;           /* 2 or 3 */
  int      basis;                /* rbfmap_basis */
  double   support_radius;       /* compact polynomials / compact TPS; optional (>0) for gaussian */
  double   shape_parameter;      /* gaussian, (inverse) multiquadrics */
  int      polynomial;           /* rbfmap_polynomial */
  int      dead_axis[3];         /* x-dead / y-dead / z-dead (global variants only) */
  unsigned vertices_per_cluster; /* PUM, default 50; clusters of more than 3900 vertices are rejected by computeMapping
                                    with RBFMAP_ERR_UNSUPPORTED (shared-memory bound of the per-cluster kernels) */
  double   relative_overlap;     /* PUM, default 0.15 */
  int      project_to_input;     /* PUM, default 1 */
  double   solver_rtol;          /* global-iterative, default 1e-9 */
  int      max_iterations;       /* global-iterative, default 1000000 (GinkgoRadialBasisFctSolver.hpp:337-374) */
  int      device_id;            /* CUDA device ordinal; -1 = current device */
  int      execution_mode;       /* rbfmap_execution_mode (PUM), default minimal-memory = the constructor default
                                    computeEvaluationOffline = false (PartitionOfUnityMapping.hpp:57) */
  int      shard_mode;           /* rbfmap_shard_mode (PUM after rbfmap_distributed_init), default DUPLICATE */
  int      rescaling;            /* rbf-global-iterative, consistent + polynomial separate: divide the interpolant by the
                                    interpolant of g(x) = 1 (PetRadialBasisFctMapping.hpp:109, 126-127, 470-490, 663-666);
                                    default 1 = the reference's hard-wired useRescaling = true */
  double   output_filter;        /* rbf-global-iterative: mapped values with |x| < output_filter become exactly 0
                                    (VecFilter(out, 1e-9), PetRadialBasisFctMapping.hpp:25-32, 675, 799); default 1e-9, 0 = off */
  int      deterministic;        /* PUM: 1 = blend by a per-vertex gather in ascending cluster order instead of atomics
                                    (bit-reproducible output), default 0 */
} rbfmap_config;

typedef struct rbfmap_handle rbfmap_handle;

/* Fills *cfg with the XML defaults (MappingConfiguration.cpp:227-248). */
void rbfmap_default_config(rbfmap_config *cfg);

/* Mapping constructor. Validates the configuration like the reference constructors do
 * (RadialBasisFctMapping.hpp:90, PartitionOfUnityMapping.hpp:165, BasisFunctions.hpp ctor checks). */
int rbfmap_create(const rbfmap_config *cfg, rbfmap_handle **out);

/* Mapping::computeMapping() (mapping/Mapping.hpp:101) with the meshes handed over by value
 * (what Mapping::setMeshes + mesh.vertex(i).rawCoords() provide). Host pointers. */
int rbfmap_compute_mapping(rbfmap_handle *h, const double *input_xyz, int64_t n_input, const double *output_xyz, int64_t n_output);

/* ---- just-in-time mapping (rbf-pum-direct only, other variants return
 * RBFMAP_ERR_UNSUPPORTED). The reference forbids it on its own GPU path (PartitionOfUnityMapping.hpp:211). ----
 *
 * Mapping::computeMapping() when the other mesh is a just-in-time mesh
 * (mesh::MeshConfiguration::getJustInTimeMappingMesh; CreateClustering.hpp:313, 469, 492): `mesh_xyz` is the mesh that
 * exists - Mapping::input() for a consistent (read) mapping, Mapping::output() for a conservative (write) mapping.
 * Host pointer. */
int rbfmap_compute_mapping_just_in_time(rbfmap_handle *h, const double *mesh_xyz, int64_t n);

/* impl::MappingDataCache (mapping/impl/MappingDataCache.hpp:23-71) on the device, one per data. */
typedef struct rbfmap_cache rbfmap_cache;

/* Mapping::initializeMappingDataCache (Mapping.hpp, PartitionOfUnityMapping.hpp:424-434): sized for the current
 * mapping and zeroed; create it after rbfmap_compute_mapping_just_in_time and again after every re-computation. */
int rbfmap_cache_create(rbfmap_handle *h, int dims, rbfmap_cache **out);
/* MappingDataCache::resetData (:60-70). */
int  rbfmap_cache_reset(rbfmap_handle *h, rbfmap_cache *cache);
void rbfmap_cache_destroy(rbfmap_cache *cache);

/* Mapping::updateMappingDataCache (PartitionOfUnityMapping.hpp:436-448): `values` = n*dims doubles on the mesh,
 * vertex-major. Consistent mappings only. */
int rbfmap_cache_update(rbfmap_handle *h, rbfmap_cache *cache, const double *values);

/* Mapping::mapConsistentAt (PartitionOfUnityMapping.hpp:450-476): evaluates the cached interpolant at `n_points`
 * arbitrary points (n_points x 3, z = 0 in 2-D); `out` = n_points*dims doubles, overwritten. A point outside every
 * cluster raises RBFMAP_ERR_UNASSIGNED with the reference's message (:314-318). */
int rbfmap_map_consistent_at(rbfmap_handle *h, rbfmap_cache *cache, const double *coords, int64_t n_points, double *out);

/* Mapping::mapConservativeAt (PartitionOfUnityMapping.hpp:382-409): adds the loads `values` (n_points*dims) given at
 * arbitrary points to the cache. Conservative mappings only. */
int rbfmap_map_conservative_at(rbfmap_handle *h, rbfmap_cache *cache, const double *coords, int64_t n_points, const double *values);

/* Mapping::completeJustInTimeMapping (PartitionOfUnityMapping.hpp:411-422): solves the cached systems and ADDS the
 * result to `out` (n*dims doubles on the mesh), like the reference accumulates into its buffer. */
int rbfmap_complete_just_in_time_mapping(rbfmap_handle *h, rbfmap_cache *cache, double *out);

/* Mesh connectivity for the scaled-consistent constraints (Mapping::scaleConsistentMapping, Mapping.cpp:175-246): the
 * elements of `mesh` (0 = Mapping::input(), 1 = Mapping::output()) as tuples of vertex IDs, vertices_per_element = 2
 * (edges: 2-D surface), 3 (triangles: 2-D volume, 3-D surface) or 4 (tetrahedra: 3-D volume) - what Mapping.cpp:190-192
 * requires for the configured constraint and dimension. Call before computeMapping; rbfmap_map() then multiplies every
 * data component of the consistent result by integral(input) / integral(output) (mesh::integrateSurface /
 * integrateVolume, mesh/Utils.cpp:11-70; factor 1 if the output integral is zero). Without connectivity rbfmap_map()
 * fails with the reference's message ("... connectivity information is missing for the mesh ..."). Host pointer. */
int rbfmap_set_connectivity(rbfmap_handle *h, int mesh, const int32_t *elements, int64_t n_elements, int vertices_per_element);

/* Page-locked host memory for the adapter's staging buffers (cudaHostAlloc / cudaFreeHost): the adapter has to pack
 * mesh.vertex(i).rawCoords() into a contiguous double[n][3] anyway (the reference's own device path does the same,
 * BatchedRBFSolver.hpp:250-271); packing into a buffer from rbfmap_host_alloc lets the host-pointer entry points copy
 * at full PCIe speed and asynchronously. Pageable pointers are accepted everywhere, just slower (driver-staged). */
int  rbfmap_host_alloc(void **ptr, int64_t bytes);
void rbfmap_host_free(void *ptr);

/* Mapping::map(Sample, VectorXd&) dispatch (mapping/Mapping.cpp:158-173): conservative ->
 * mapConservative, otherwise mapConsistent. `in` has n_input*dims, `out` n_output*dims doubles.
 * `out` is overwritten with the mapped values (the reference accumulates into a pre-zeroed vector,
 * precice/impl/DataContext.cpp:148-150). Scaled-consistent constraints: mapConsistent followed by the
 * post-scaling of Mapping.cpp:175-246 (needs rbfmap_set_connectivity). Host pointers. */
int rbfmap_map(rbfmap_handle *h, const double *in, int dims, double *out);
int rbfmap_map_consistent(rbfmap_handle *h, const double *in, int dims, double *out);   /* Mapping.hpp:337 */
int rbfmap_map_conservative(rbfmap_handle *h, const double *in, int dims, double *out); /* Mapping.hpp:324 */

/* rbfmap_map for a sharded mapping (rbfmap_distributed_init) whose ranks each want THEIR part of the result - the mirror of
 * the upload, where every rank sends 1/world of the values over PCIe: the partial results are reduce-scattered over NVLink
 * (nothing to exchange where every rank already holds the complete result) and only the values of the output vertices
 * [*begin, *end) = rbfmap_partition_rows(n_output, world, rank) are copied to `out[*begin * dims .. *end * dims)`; the rest
 * of `out` (a full-size array, like `in`) is not touched. On a single device the slice is the whole result. Host pointers. */
int rbfmap_map_sliced(rbfmap_handle *h, const double *in, int dims, double *out, int64_t *begin, int64_t *end);

/* Several samples in one call: DataContext::mapData (precice/impl/DataContext.cpp:110-171) maps every time stample of a
 * data, and every data of a mesh, with a separate Mapping::map(); handing them over together lets the device treat them as
 * the components of one field (one gather, one factor load and one basis-function evaluation per cluster for up to three
 * components). Sample k has dims[k] components, in[k] = n_input * dims[k] and out[k] = n_output * dims[k] doubles; the
 * results equal those of `count` separate rbfmap_map() calls. Host pointers. */
int rbfmap_map_batch(rbfmap_handle *h, int count, const double *const *in, const int *dims, double *const *out);

/* Mapping::map(Sample, VectorXd&, VectorXd &lastSolution) (Mapping.cpp:150-156): the variant for mappings with
 * Mapping::requiresInitialGuess() (InitialGuessRequirement::Required, PetRadialBasisFctMapping.hpp:176) - the caller
 * (precice/impl/DataContext.cpp:159-161) owns one `guess` vector per data and hands it to every map() of that data: the
 * iterative solve starts from it and stores its solution there (loadInitialGuessForDim / storeInitialGuessForDim,
 * PetRadialBasisFctMapping.hpp:526-561). `guess` holds rbfmap_initial_guess_size(h) * dims doubles, dimension-major;
 * zero-fill it before the first use and after every computeMapping. Mappings without iterative solver report size 0
 * and ignore `guess`. rbfmap_map() on an iterative mapping keeps ONE internal guess per handle instead (the behaviour of
 * GinkgoRadialBasisFctSolver.hpp:218-221, 437). Host pointers. */
int64_t rbfmap_initial_guess_size(const rbfmap_handle *h);
int     rbfmap_map_with_initial_guess(rbfmap_handle *h, const double *in, int dims, double *out, double *guess);

/* Same two calls with DEVICE pointers (coordinates / values already resident in HBM); used by
 * device-resident callers and by bench.py's kernel-only figure. The work is enqueued on the
 * library's stream and the call returns after it completed. The library's stream is not
 * ordered behind the caller's streams: the device buffers must be completely written (the
 * producing stream synchronised or joined by an event the caller waited for) before the call.
 * rbfmap_compute_mapping_device reads the meshes in place (no copy is kept: the mapping owns only what it derived from
 * them, so the buffers may be freed or overwritten as soon as the call has returned). */
int rbfmap_compute_mapping_device(rbfmap_handle *h, const double *d_input_xyz, int64_t n_input, const double *d_output_xyz, int64_t n_output);
int rbfmap_map_device(rbfmap_handle *h, const double *d_in, int dims, double *d_out);

/* Multi-GPU, one process per GPU: ONE mapping sharded over the ranks of a group. Every rank is given the
 * same (replicated) meshes and values and makes the same sequence of calls; rank 0 obtains the 128-byte NCCL id, the host
 * application broadcasts it (MPI / sockets / torch.distributed), then every rank calls rbfmap_distributed_init before
 * computeMapping. With host pointers every rank uploads only its 1/world slice over PCIe and the library all-gathers the
 * slices over NVLink.
 *  - rbf-pum-direct: the lattice of cluster centres is that of the whole mesh (bounding box and cluster radius are
 *    computed by every rank from the complete input mesh), so the cluster set is the one of the single-device mapping.
 *    The lattice is cut into slabs along its longest axis, balanced by the vertex density; every rank bins, clusters,
 *    factorises and maps only its own slab plus a halo of a few cluster radii
 *    (rbfmap_config.shard_mode): DUPLICATE - a rank owns a slab of output vertices and processes every cluster that
 *    reaches into it, map() needs no collective and fills only the owned vertices of `out` (the others are zero; the
 *    complete result is the sum over the ranks; rbfmap_stats.shard_* describe the slab); EXCHANGE (always used for
 *    conservative constraints) - every cluster has one owner, the weight sums (computeMapping) and the partial blends
 *    (map) are all-reduced and every rank returns the complete result. The reference instead maps rank-local partitions
 *    independently (BatchedRBFSolver.hpp:29-31, halo PartitionOfUnityMapping.hpp:479-533), which changes the clustering
 *    with the partitioning; here the result does not depend on the number of ranks (1e-10).
 *  - rbf-global-iterative shards the rows of its matrix-free CG operator: a rank owns a contiguous slice of the
 *    cell-ordered rows, the direction vector is all-gathered and the dot products all-reduced (the role of PETSc's
 *    distributed KSP in PetRadialBasisFctMapping.hpp:246-263, 629-633).
 *  - rbf-global-direct stays on one device (RBFMAP_ERR_UNSUPPORTED). */
int rbfmap_nccl_unique_id(char id[128]);
/* id == NULL joins a group WITHOUT communicator: accepted by rbf-pum-direct for consistent constraints in duplicate mode
 * only, where no collective is needed (the ranks may then even be handles of one process); every rank uploads complete
 * meshes and raises its own errors. */
int rbfmap_distributed_init(rbfmap_handle *h, const char id[128], int rank, int world);
/* The contiguous slice [*begin, *end) of n rows owned by `rank` (host-side helper, no device needed). */
int rbfmap_partition_rows(int64_t n, int world, int rank, int64_t *begin, int64_t *end);
/* Host-side helper of the sharded rbf-pum-direct (no device needed): given the histogram of the cluster centres along
 * the slab axis, the cut positions 0 = cuts[0] <= ... <= cuts[world] = n_bins that minimise the largest number of
 * centres one rank processes, when rank r processes the bins [cuts[r] - halo_bins, cuts[r+1] + halo_bins). */
int rbfmap_partition_slabs(const int64_t *histogram, int n_bins, int world, int halo_bins, int *cuts);

/* Re-partitioning hooks of a parallel participant (Mapping::tagMeshFirstRound / tagMeshSecondRound, Mapping.hpp:179-182,
 * called by partition/ReceivedPartition.cpp:921-954): which vertices of the received (remote) mesh does this rank keep?
 * The remote mesh is Mapping::input() for consistent and Mapping::output() for conservative constraints; `tags` has one
 * byte per vertex of THAT mesh, the call sets tags[i] = 1 for the vertices to keep and leaves the others untouched
 * (Vertex::tag()). Host pointers.
 *  - global variants (RadialBasisFctBaseMapping.hpp:124-190): bounding box of the local mesh, expanded by the support
 *    radius (all vertices if the basis function has global support); second round: box around the owned vertices
 *    (`owner[i] != 0`) of the remote mesh, expanded by the support radius.
 *  - rbf-pum-direct (PartitionOfUnityMapping.hpp:479-539): bounding box of the local mesh, expanded by 2 x the cluster
 *    radius, which is estimated on the unfiltered remote mesh (rbfmap_estimate_cluster_radius) unless a computed mapping
 *    already knows it; the second round does nothing. */
int rbfmap_tag_mesh_first_round(rbfmap_handle *h, const double *input_xyz, int64_t n_input, const double *output_xyz, int64_t n_output,
                                unsigned char *tags);
int rbfmap_tag_mesh_second_round(rbfmap_handle *h, const double *mesh_xyz, int64_t n, const unsigned char *owner, unsigned char *tags);
/* impl::estimateClusterRadius(verticesPerCluster, mesh, boundingBox) (impl/CreateClustering.hpp:223-278): median over
 * 1 + 2 dim probe points of the distance to the vertices-per-cluster-th nearest vertex. rbf-pum-direct only. */
int rbfmap_estimate_cluster_radius(rbfmap_handle *h, const double *mesh_xyz, int64_t n, const double bb_min[3], const double bb_max[3],
                                   double *radius);

/* Mapping::clear() (Mapping.hpp:134), Mapping::hasComputedMapping(), destructor, Mapping::getName(). */
int         rbfmap_clear(rbfmap_handle *h);
int         rbfmap_has_computed_mapping(const rbfmap_handle *h);
void        rbfmap_destroy(rbfmap_handle *h);
const char *rbfmap_get_name(const rbfmap_handle *h);

/* Message of the last failing call on this handle (or of a failed rbfmap_create if h == NULL). */
const char *rbfmap_last_error(const rbfmap_handle *h);

/* Counters for profiling::Event::addData ("n clusters", PartitionOfUnityMapping.hpp:203) and for the
 * roofline bookkeeping in bench.py (algorithmic flops per SURVEY.md §8d). */
typedef struct rbfmap_stats {
  int64_t n_in;               /* solver-side inMesh size (after the conservative swap) */
  int64_t n_out;              /* solver-side outMesh size */
  int64_t n_clusters;         /* PUM */
  double  cluster_radius;     /* PUM */
  int64_t sum_n;              /* sum over clusters of in-vertices n_c */
  int64_t sum_m;              /* sum over clusters of out-vertices m_c */
  int64_t sum_n2;             /* sum n_c^2 */
  int64_t sum_n3;             /* sum n_c^3 */
  int64_t sum_mn;             /* sum m_c n_c */
  int64_t max_n;              /* largest cluster */
  int64_t device_bytes;       /* bytes held on the device by this mapping */
  int     iterations;         /* global-iterative: iterations of the last solve (max over components) */
  double  residual;           /* global-iterative: relative residual of the last solve */
  float   ms_compute_kernels; /* device time of the last computeMapping (CUDA events, whole call) */
  float   ms_map_kernels;     /* device time of the last map */
  float   ms_assembly_factor; /* device time of the dominant assembly(+factorisation) kernel(s) of the last computeMapping */
  int     kernel_launches;    /* kernels launched by the last compute + map calls */
  /* sharded rbf-pum-direct (rbfmap_distributed_init): n_clusters and the sums above describe what THIS rank processes */
  int64_t n_clusters_global;  /* clusters of the whole mapping (-1: unknown, a sharded mapping joined without communicator) */
  int     shard_axis;         /* axis of the slab decomposition; -1 = not sharded */
  double  shard_lo, shard_hi; /* this rank owns lo <= x[axis] < hi (output vertices in duplicate mode, cluster centres in exchange mode) */
  float   ms_replicated;      /* device time of the phases of computeMapping every rank repeats (bounds, radius, binning, clustering) */
  float   ms_collectives;     /* device time of the NCCL collectives of the last map */
  /* execution_mode = minimal-compute */
  int64_t evaluation_bytes;   /* bytes of the stored per-cluster evaluation operators */
  float   ms_evaluation_offline; /* device time spent building them in the last computeMapping */
} rbfmap_stats;
int rbfmap_get_stats(const rbfmap_handle *h, rbfmap_stats *stats);

/* Debug/inspection access used by the parity tests (sizes first, then data). PUM only.
 * ids are vertex IDs of the solver-side inMesh / outMesh, sorted ascending per cluster
 * (impl/SphericalVertexCluster.hpp:161-162). Any output pointer may be NULL. */
int rbfmap_pum_get_clusters(const rbfmap_handle *h, double *centers /* n_clusters*3 */, int64_t *in_offsets /* n_clusters+1 */,
                            int64_t *out_offsets /* n_clusters+1 */);
int rbfmap_pum_get_cluster_ids(const rbfmap_handle *h, int32_t *in_ids /* sum_n */, int32_t *out_ids /* sum_m */);

/* Basis function evaluated on the device for `n` radii (host pointers) — exposes a1 of SURVEY §8
 * (mapping/impl/BasisFunctions.hpp) for bit-exact parity tests. */
int rbfmap_eval_basis(int basis, double support_radius, double shape_parameter, const double *radii, int64_t n, double *values);

/* Test hook: phi(|a_i - b_i|) for n vertex pairs (n x 3 each, z ignored for dimensions = 2) computed by the PAIR
 * evaluation of the per-cluster algebra kernels (pairBasis / pairBasisBatch, precice_b200/csrc/pum_device.cuh): for the
 * compact polynomials fused multiply-adds, a reciprocal-root seed with a second-order correction and an integer clamp at
 * the support, a few ulp of phi(0) away from the operation-by-operation functors that rbfmap_eval_basis exposes;
 * batched != 0 selects the four-wide staged form. Host pointers. */
int rbfmap_debug_pair_basis(int basis, double support_radius, double shape_parameter, int dimensions, const double *a, const double *b, int64_t n,
                            int batched, double *values);

/* Test hook: the branch-free square root the kernels apply to squared vertex distances (host pointers). */
int rbfmap_debug_dist_sqrt(const double *x, int64_t n, double *out);

/* sizeof(rbfmap_config) and sizeof(rbfmap_stats) as compiled into the library: lets a binding written in another
 * language (ctypes, cgo, JNI) verify its mirror of the two structs at load time. */
void rbfmap_struct_sizes(int *config_bytes, int *stats_bytes);

/* Library / device probe: number of visible CUDA devices (0 if none), compute capability of device 0. */
int rbfmap_device_count(void);
int rbfmap_device_info(int device, int *sm_major, int *sm_minor, int *sm_count, int64_t *global_mem_bytes);

/* Measured FP64 FMA peak of the current device (dependency-free DFMA loop, best of 4): the roofline denominator of
 * the FP64-bound kernels; bench.py reports it next to every roofline fraction. */
int rbfmap_measure_fp64_peak(double *tflops, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* RBFMAP_H */
