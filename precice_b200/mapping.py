"""Host-side mirror of the reference's mapping classes for the RBF hot path, on top of the C ABI.

Class and argument names follow the reference so that the parity tests read like the reference's own tests:
  PartitionOfUnityMapping  <- mapping/PartitionOfUnityMapping.hpp:48-57
  RadialBasisFctMapping    <- mapping/RadialBasisFctMapping.hpp:40-46 (global-direct / global-iterative)
Methods: computeMapping / map / mapConsistent / mapConservative / clear / hasComputedMapping / getName
(mapping/Mapping.hpp:101, :134, :197, :324, :337; dispatch mapping/Mapping.cpp:158-173). Snake-case aliases are
provided for the shared test tables (compute_mapping, map).

All arithmetic happens in librbfmap.so on the GPU; this file only marshals numpy / torch buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BASIS, CONSTRAINT, POLYNOMIAL, VARIANT, Config, Stats


class MappingError(RuntimeError):
    """precice::Error equivalent (precice/Exceptions.hpp:9-14); .status holds the rbfmap_status name."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code
        self.status = _lib.STATUS.get(code, str(code))


def as_xyz(coords) -> np.ndarray:
    """(N, dim) -> C-contiguous (N, 3) float64 with z = 0 for 2-D meshes (mesh/Vertex.hpp:75-77)."""
    a = np.asarray(coords, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 3)
    if a.shape[1] == 2:
        a = np.concatenate([a, np.zeros((a.shape[0], 1))], axis=1)
    return np.ascontiguousarray(a)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _sync_producer(t):
    """Device-pointer entry points: the tensors must be complete before the library's own stream reads them."""
    import torch
    torch.cuda.current_stream(t.device).synchronize()


class _Mapping:
    def __init__(self, cfg: Config):
        self._L = _lib.lib()
        self._cfg = cfg
        h = C.c_void_p()
        rc = self._L.rbfmap_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise MappingError(rc, self._L.rbfmap_last_error(None).decode())
        self._h = h
        self._n_input = self._n_output = 0

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise MappingError(rc, self._L.rbfmap_last_error(self._h).decode())

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.rbfmap_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- mapping::Mapping interface
    def computeMapping(self, input_xyz, output_xyz):
        """Mapping::computeMapping() after Mapping::setMeshes(input, output). Accepts numpy (host) or CUDA torch tensors."""
        if _is_torch(input_xyz) and input_xyz.is_cuda:
            a, b = input_xyz.contiguous(), output_xyz.contiguous()
            assert a.dtype.is_floating_point and a.element_size() == 8 and a.shape[-1] == 3 and b.shape[-1] == 3
            self._n_input, self._n_output = a.shape[0], b.shape[0]
            _sync_producer(a)  # the library's stream is not ordered behind torch's (include/rbfmap.h)
            self._check(self._L.rbfmap_compute_mapping_device(self._h, a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0]))
            return
        if _is_torch(input_xyz):  # pinned / pageable host tensors
            a, b = input_xyz.contiguous(), output_xyz.contiguous()
            self._n_input, self._n_output = a.shape[0], b.shape[0]
            self._check(self._L.rbfmap_compute_mapping(self._h, a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0]))
            return
        a, b = as_xyz(input_xyz), as_xyz(output_xyz)
        self._n_input, self._n_output = a.shape[0], b.shape[0]
        self._check(self._L.rbfmap_compute_mapping(self._h, a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0]))

    def _map(self, fn_host, values, dims, out=None):
        if not self.hasComputedMapping():
            self._check(fn_host(self._h, None, dims, None))  # -> RBFMAP_ERR_STATE with the library's message
        if _is_torch(values):
            import torch
            v = values.contiguous()
            assert v.numel() == self._n_input * dims, (v.numel(), self._n_input, dims)
            if v.is_cuda:
                if out is None:
                    out = torch.empty(self._n_output * dims, dtype=torch.float64, device=v.device)
                _sync_producer(v)
                self._check(self._L.rbfmap_map_device(self._h, v.data_ptr(), dims, out.data_ptr()))
                return out
            if out is None:
                out = torch.empty(self._n_output * dims, dtype=torch.float64)
            self._check(fn_host(self._h, v.data_ptr(), dims, out.data_ptr()))
            return out
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
        assert v.size == self._n_input * dims, (v.size, self._n_input, dims)
        if out is None:
            out = np.zeros(self._n_output * dims)
        self._check(fn_host(self._h, v.ctypes.data, dims, out.ctypes.data))
        return out

    def map(self, values, dims=1, out=None):
        """Mapping::map(Sample, VectorXd&): dispatch on the constraint (Mapping.cpp:158-173)."""
        return self._map(self._L.rbfmap_map, values, dims, out)

    def mapSliced(self, values, dims=1, out=None):
        """rbfmap_map_sliced: map() on a sharded mapping, every rank reads back only its slice of the output vertices.
        Returns (out, begin, end): `out` is a full-size host array of which [begin * dims, end * dims) has been written."""
        b, e = C.c_int64(0), C.c_int64(0)
        if _is_torch(values):
            import torch
            v = values.contiguous()
            assert not v.is_cuda and v.numel() == self._n_input * dims
            if out is None:
                out = torch.zeros(self._n_output * dims, dtype=torch.float64)
            self._check(self._L.rbfmap_map_sliced(self._h, v.data_ptr(), dims, out.data_ptr(), C.byref(b), C.byref(e)))
            return out, b.value, e.value
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
        assert v.size == self._n_input * dims, (v.size, self._n_input, dims)
        if out is None:
            out = np.zeros(self._n_output * dims)
        self._check(self._L.rbfmap_map_sliced(self._h, v.ctypes.data, dims, out.ctypes.data, C.byref(b), C.byref(e)))
        return out, b.value, e.value

    def mapBatch(self, samples, dims):
        """Maps several samples (time stamples / data of one mesh, DataContext.cpp:110-171) in one call; returns one array per
        sample, equal to [self.map(s, d) for s, d in zip(samples, dims)]."""
        ins = [np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1)) for v in samples]
        for v, d in zip(ins, dims):
            assert v.size == self._n_input * d
        outs = [np.zeros(self._n_output * d) for d in dims]
        k = len(ins)
        pin = (C.c_void_p * k)(*[v.ctypes.data for v in ins])
        pout = (C.c_void_p * k)(*[v.ctypes.data for v in outs])
        pd = (C.c_int * k)(*[int(d) for d in dims])
        self._check(self._L.rbfmap_map_batch(self._h, k, pin, pd, pout))
        return outs

    def initialGuessSize(self) -> int:
        """Per-dimension length of the caller-owned initial guess (0: the mapping needs none, Mapping::requiresInitialGuess)."""
        return int(self._L.rbfmap_initial_guess_size(self._h))

    def mapWithInitialGuess(self, values, dims, guess, out=None):
        """Mapping::map(input, output, lastSolution) (Mapping.cpp:150-156): `guess` (float64, initialGuessSize() * dims,
        zero before the first use) is read as the start of the iterative solve and overwritten with its solution."""
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
        assert v.size == self._n_input * dims
        assert guess.dtype == np.float64 and guess.flags.c_contiguous and guess.size == self.initialGuessSize() * dims
        if out is None:
            out = np.zeros(self._n_output * dims)
        self._check(self._L.rbfmap_map_with_initial_guess(self._h, v.ctypes.data, dims, out.ctypes.data, guess.ctypes.data))
        return out

    def mapConsistent(self, values, dims=1, out=None):
        return self._map(self._L.rbfmap_map_consistent, values, dims, out)

    def mapConservative(self, values, dims=1, out=None):
        return self._map(self._L.rbfmap_map_conservative, values, dims, out)

    # ---- just-in-time mapping (Mapping.hpp: initializeMappingDataCache, updateMappingDataCache, mapConsistentAt,
    # mapConservativeAt, completeJustInTimeMapping); rbf-pum-direct only ----
    def computeMappingJustInTime(self, mesh_xyz):
        """computeMapping() when the other mesh is a just-in-time mesh: mesh_xyz = input() (consistent) / output() (conservative)."""
        a = as_xyz(mesh_xyz)
        self._n_mesh = a.shape[0]
        self._n_input, self._n_output = (a.shape[0], 0) if self._cfg.constraint != 1 else (0, a.shape[0])
        self._check(self._L.rbfmap_compute_mapping_just_in_time(self._h, a.ctypes.data, a.shape[0]))

    def initializeMappingDataCache(self, dims):
        c = C.c_void_p()
        self._check(self._L.rbfmap_cache_create(self._h, dims, C.byref(c)))
        return MappingDataCache(self, c, dims)

    def updateMappingDataCache(self, cache, values):
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
        assert v.size == self._n_mesh * cache.dims, (v.size, self._n_mesh, cache.dims)
        self._check(self._L.rbfmap_cache_update(self._h, cache._c, v.ctypes.data))

    def mapConsistentAt(self, coords, cache) -> np.ndarray:
        x = as_xyz(coords)
        out = np.zeros(x.shape[0] * cache.dims)
        self._check(self._L.rbfmap_map_consistent_at(self._h, cache._c, x.ctypes.data, x.shape[0], out.ctypes.data))
        return out

    def mapConservativeAt(self, coords, source, cache):
        x = as_xyz(coords)
        v = np.ascontiguousarray(np.asarray(source, dtype=np.float64).reshape(-1))
        assert v.size == x.shape[0] * cache.dims
        self._check(self._L.rbfmap_map_conservative_at(self._h, cache._c, x.ctypes.data, x.shape[0], v.ctypes.data))

    def completeJustInTimeMapping(self, cache, out):
        assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous and out.size == self._n_mesh * cache.dims
        self._check(self._L.rbfmap_complete_just_in_time_mapping(self._h, cache._c, out.ctypes.data))

    def setConnectivity(self, mesh: str, elements):
        """Edges / triangles / tetrahedra (rows of vertex IDs) of Mapping::input() ("input") or output() ("output") for the
        scaled-consistent constraints (Mapping.cpp:175-246); call before computeMapping."""
        e = np.ascontiguousarray(elements, dtype=np.int32)
        assert e.ndim == 2 and e.shape[1] in (2, 3, 4)
        self._check(self._L.rbfmap_set_connectivity(self._h, {"input": 0, "output": 1}[mesh], e.ctypes.data, e.shape[0], e.shape[1]))

    # -- re-partitioning hooks (Mapping.hpp:179-182)
    def tagMeshFirstRound(self, input_xyz, output_xyz, tags=None) -> np.ndarray:
        """Tags (1) the vertices of the remote mesh - input() for consistent, output() for conservative constraints - that
        this rank needs; `tags` (uint8, one per vertex of the remote mesh) is updated in place and returned."""
        a, b = as_xyz(input_xyz), as_xyz(output_xyz)
        n_remote = b.shape[0] if self._cfg.constraint == 1 else a.shape[0]
        if tags is None:
            tags = np.zeros(n_remote, dtype=np.uint8)
        assert tags.dtype == np.uint8 and tags.size == n_remote and tags.flags.c_contiguous
        self._check(self._L.rbfmap_tag_mesh_first_round(self._h, a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], tags.ctypes.data))
        return tags

    def tagMeshSecondRound(self, remote_xyz, owner, tags) -> np.ndarray:
        a = as_xyz(remote_xyz)
        owner = np.ascontiguousarray(owner, dtype=np.uint8)
        assert tags.dtype == np.uint8 and tags.size == a.shape[0] == owner.size
        self._check(self._L.rbfmap_tag_mesh_second_round(self._h, a.ctypes.data, a.shape[0], owner.ctypes.data, tags.ctypes.data))
        return tags

    def estimateClusterRadius(self, mesh_xyz, bb_min, bb_max) -> float:
        """impl::estimateClusterRadius(verticesPerCluster, mesh, bb) (impl/CreateClustering.hpp:223-278); rbf-pum-direct."""
        a = as_xyz(mesh_xyz)
        lo = (C.c_double * 3)(*[float(v) for v in list(bb_min) + [0.0] * (3 - len(bb_min))])
        hi = (C.c_double * 3)(*[float(v) for v in list(bb_max) + [0.0] * (3 - len(bb_max))])
        r = C.c_double(0)
        self._check(self._L.rbfmap_estimate_cluster_radius(self._h, a.ctypes.data, a.shape[0], lo, hi, C.byref(r)))
        return r.value

    def clear(self):
        self._check(self._L.rbfmap_clear(self._h))

    def hasComputedMapping(self) -> bool:
        return bool(self._L.rbfmap_has_computed_mapping(self._h))

    def getName(self) -> str:
        return self._L.rbfmap_get_name(self._h).decode()

    # -- multi-GPU: one mapping sharded over a group of ranks (see include/rbfmap.h)
    def distributedInit(self, unique_id, rank: int, world: int):
        """unique_id: the 128 bytes of nccl_unique_id() (broadcast from rank 0), or None for a group without communicator
        (rbf-pum-direct, consistent, duplicate mode only: no collective is needed)."""
        assert unique_id is None or len(unique_id) == 128
        self._check(self._L.rbfmap_distributed_init(self._h, unique_id, rank, world))

    # -- extras
    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.rbfmap_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    # snake-case aliases used by the shared test tables
    compute_mapping = computeMapping


class MappingDataCache:
    """impl::MappingDataCache (mapping/impl/MappingDataCache.hpp:23-71): device-resident per-cluster coefficients of one data."""

    def __init__(self, mapping, handle, dims):
        self._m, self._c, self.dims = mapping, handle, dims

    def resetData(self):
        self._m._check(self._m._L.rbfmap_cache_reset(self._m._h, self._c))

    def __del__(self):
        try:
            self._m._L.rbfmap_cache_destroy(self._c)
        except Exception:
            pass


def _config(variant, constraint, dim, basis, param, polynomial, dead_axis=(False, False, False), vpc=50, overlap=0.15,
            project=True, gauss_support=None, solver_rtol=1e-9, max_iterations=1000000, device_id=-1,
            execution_mode="minimal-memory", shard_mode="duplicate", deterministic=False, rescaling=True,
            output_filter=1e-9) -> Config:
    cfg = Config()
    _lib.lib().rbfmap_default_config(C.byref(cfg))
    cfg.variant = VARIANT[variant]
    cfg.constraint = CONSTRAINT[constraint]
    cfg.dimensions = dim
    cfg.basis = BASIS[basis]
    # the single functor argument: support radius (compact functions) or shape parameter (gaussian, (i)mq)
    if basis in ("gaussian", "multiquadrics", "inverse-multiquadrics"):
        cfg.shape_parameter = float(param)
        cfg.support_radius = float(gauss_support) if gauss_support else 0.0
    else:
        cfg.support_radius = float(param)
    cfg.polynomial = POLYNOMIAL[polynomial]
    for d in range(3):
        cfg.dead_axis[d] = int(bool(dead_axis[d]))
    cfg.vertices_per_cluster = int(vpc)
    cfg.relative_overlap = float(overlap)
    cfg.project_to_input = int(bool(project))
    cfg.solver_rtol = float(solver_rtol)
    cfg.max_iterations = int(max_iterations)
    cfg.device_id = int(device_id)
    cfg.execution_mode = _lib.EXECUTION_MODE[execution_mode]
    cfg.shard_mode = _lib.SHARD_MODE[shard_mode]
    cfg.deterministic = int(bool(deterministic))
    cfg.rescaling = int(bool(rescaling))
    cfg.output_filter = float(output_filter)
    return cfg


class PartitionOfUnityMapping(_Mapping):
    """mapping::PartitionOfUnityMapping<RBF>(constraint, dimension, function, polynomial, verticesPerCluster,
    relativeOverlap, projectToInput) — PartitionOfUnityMapping.hpp:48-57, XML tag <mapping:rbf-pum-direct>."""

    def __init__(self, constraint, dim, basis, param, polynomial="separate", vertices_per_cluster=50,
                 relative_overlap=0.15, project_to_input=True, gauss_support=None, device_id=-1,
                 execution_mode="minimal-memory", shard_mode="duplicate", deterministic=False):
        """execution_mode: <executor execution-mode=...> / computeEvaluationOffline (PartitionOfUnityMapping.hpp:57);
        shard_mode, deterministic: see include/rbfmap.h."""
        super().__init__(_config("rbf-pum-direct", constraint, dim, basis, param, polynomial, vpc=vertices_per_cluster,
                                 overlap=relative_overlap, project=project_to_input, gauss_support=gauss_support,
                                 device_id=device_id, execution_mode=execution_mode, shard_mode=shard_mode,
                                 deterministic=deterministic))

    def clusters(self):
        """(radius, centers[nc,3], in_offsets, out_offsets, in_ids, out_ids) — inspection for the parity tests."""
        st = self.stats()
        nc = st["n_clusters"]
        centers = np.zeros((nc, 3))
        in_off = np.zeros(nc + 1, dtype=np.int64)
        out_off = np.zeros(nc + 1, dtype=np.int64)
        self._check(self._L.rbfmap_pum_get_clusters(self._h, centers.ctypes.data, in_off.ctypes.data, out_off.ctypes.data))
        in_ids = np.zeros(max(int(st["sum_n"]), 1), dtype=np.int32)
        out_ids = np.zeros(max(int(st["sum_m"]), 1), dtype=np.int32)
        self._check(self._L.rbfmap_pum_get_cluster_ids(self._h, in_ids.ctypes.data, out_ids.ctypes.data))
        return st["cluster_radius"], centers, in_off, out_off, in_ids[:st["sum_n"]], out_ids[:st["sum_m"]]


class RadialBasisFctMapping(_Mapping):
    """mapping::RadialBasisFctMapping<SOLVER>(constraint, dimension, function, deadAxis, polynomial) —
    RadialBasisFctMapping.hpp:40-46; solver = "direct" (<mapping:rbf-global-direct>) or "iterative"
    (<mapping:rbf-global-iterative>, matrix-free CG)."""

    def __init__(self, constraint, dim, basis, param=0.0, dead_axis=(False, False, False), polynomial="separate",
                 gauss_support=None, solver="direct", solver_rtol=1e-9, max_iterations=1000000, device_id=-1,
                 rescaling=True, output_filter=1e-9):
        """rescaling / output_filter: semantics of the reference's iterative path (PetRadialBasisFctMapping.hpp:126-127,
        470-490, 663-666 and VecFilter :675, :799); ignored by the direct solver."""
        variant = "rbf-global-direct" if solver == "direct" else "rbf-global-iterative"
        super().__init__(_config(variant, constraint, dim, basis, param, polynomial, dead_axis=dead_axis,
                                 gauss_support=gauss_support, solver_rtol=solver_rtol, max_iterations=max_iterations,
                                 device_id=device_id, rescaling=rescaling, output_filter=output_filter))


def eval_basis(basis: str, param: float, radii, gauss_support=None) -> np.ndarray:
    """Evaluates a basis function on the device (impl/BasisFunctions.hpp functors)."""
    L = _lib.lib()
    r = np.ascontiguousarray(np.atleast_1d(np.asarray(radii, dtype=np.float64)))
    out = np.empty_like(r)
    if basis in ("gaussian", "multiquadrics", "inverse-multiquadrics"):
        support, shape = (float(gauss_support) if gauss_support else 0.0), float(param)
    else:
        support, shape = float(param), 0.0
    rc = L.rbfmap_eval_basis(BASIS[basis], support, shape, r.ctypes.data_as(C.POINTER(C.c_double)), r.size,
                             out.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise MappingError(rc, L.rbfmap_last_error(None).decode())
    return out


def nccl_unique_id() -> bytes:
    """128-byte NCCL id created on the calling rank; broadcast it to the other ranks before distributedInit."""
    L = _lib.lib()
    buf = C.create_string_buffer(128)
    rc = L.rbfmap_nccl_unique_id(buf)
    if rc != 0:
        raise MappingError(rc, L.rbfmap_last_error(None).decode())
    return buf.raw


def partition_rows(n: int, world: int, rank: int):
    """Row slice [begin, end) of `rank` in the sharded CG (host-side helper)."""
    b, e = C.c_int64(0), C.c_int64(0)
    rc = _lib.lib().rbfmap_partition_rows(n, world, rank, C.byref(b), C.byref(e))
    if rc != 0:
        raise MappingError(rc, "invalid partition request")
    return b.value, e.value


def device_count() -> int:
    return _lib.lib().rbfmap_device_count()
