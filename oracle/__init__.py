"""CPU ORACLE — test infrastructure, NOT the product.

ctypes front-end of ``oracle/rbf_oracle*.{hpp,cpp}``: a dependency-free C++ restatement of the
algorithm of preCICE's RBF data-mapping hot path (reference file:line citations are in the C++
sources). Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; ``precice_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librbf_oracle.so")
_SOURCES = ["rbf_oracle_capi.cpp", "rbf_oracle.hpp", "rbf_oracle_mapping.hpp", "Makefile"]

# numeric values follow mapping/config/MappingConfigurationTypes.hpp:11-30 and Mapping.hpp:30-35
BASIS = {
    "compact-polynomial-c0": 0, "compact-polynomial-c2": 1, "compact-polynomial-c4": 2,
    "compact-polynomial-c6": 3, "compact-polynomial-c8": 4, "thin-plate-splines": 5,
    "multiquadrics": 6, "inverse-multiquadrics": 7, "volume-splines": 8, "gaussian": 9,
    "compact-tps-c2": 10,
}
POLYNOMIAL = {"on": 0, "off": 1, "separate": 2}
CONSTRAINT = {"consistent": 0, "conservative": 1, "scaled-consistent-surface": 2, "scaled-consistent-volume": 3}


class OracleError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compiles the oracle with the committed Makefile (g++ only, a few seconds)."""
    stale = force or not os.path.exists(_LIB_PATH)
    if not stale:
        t = os.path.getmtime(_LIB_PATH)
        stale = any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SOURCES)
    if stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_LIB_PATH)
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_pum_create.restype = C.c_void_p
        _lib.orc_pum_cache_create.restype = C.c_void_p
        _lib.orc_global_create.restype = C.c_void_p
        _lib.orc_pum_radius.restype = C.c_double
        for name in ("orc_pum_compute", "orc_pum_compute_sampled", "orc_pum_map", "orc_pum_clear", "orc_pum_destroy", "orc_pum_radius",
                     "orc_pum_num_clusters", "orc_pum_cluster_sizes", "orc_pum_cluster_data", "orc_global_compute",
                     "orc_global_map", "orc_global_clear", "orc_global_destroy"):
            getattr(_lib, name).argtypes = None
    return _lib


def _check(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def as_xyz(coords) -> np.ndarray:
    """(N, dim) -> contiguous (N, 3) float64 with z = 0 for 2-D meshes (mesh/Vertex.hpp:75-77)."""
    a = np.asarray(coords, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 3)
    if a.shape[1] == 2:
        a = np.concatenate([a, np.zeros((a.shape[0], 1))], axis=1)
    return np.ascontiguousarray(a)


def _gauss_support(s):
    return float("nan") if s is None else float(s)


def basis_eval(basis: str, param: float, r, gauss_support=None) -> np.ndarray:
    r = np.ascontiguousarray(np.atleast_1d(np.asarray(r, dtype=np.float64)))
    out = np.empty_like(r)
    _check(lib().orc_basis_eval(C.c_int(BASIS[basis]), C.c_double(param), C.c_double(_gauss_support(gauss_support)),
                                _dp(r), C.c_int(r.size), _dp(out)))
    return out


def basis_params(basis: str, param: float, gauss_support=None):
    p = np.zeros(3)
    spd = C.c_int(0)
    sup = C.c_double(0)
    _check(lib().orc_basis_params(C.c_int(BASIS[basis]), C.c_double(param), C.c_double(_gauss_support(gauss_support)),
                                  _dp(p), C.byref(spd), C.byref(sup)))
    return p, bool(spd.value), sup.value


def create_clustering(in_xyz, out_xyz, dim, overlap, vpc, project):
    a, b = as_xyz(in_xyz), as_xyz(out_xyz)
    radius = C.c_double(0)
    count = C.c_int(0)
    cap = max(16, 64 * (a.shape[0] // max(1, vpc) + 64))
    centers = np.zeros((cap, 3))
    _check(lib().orc_create_clustering(_dp(a), C.c_int(a.shape[0]), _dp(b), C.c_int(b.shape[0]), C.c_int(dim),
                                       C.c_double(overlap), C.c_uint(vpc), C.c_int(int(project)), C.byref(radius),
                                       _dp(centers), C.c_int(cap), C.byref(count)))
    return radius.value, centers[:min(cap, count.value)].copy(), count.value


def estimate_cluster_radius(xyz, dim, vpc, bb_min, bb_max) -> float:
    """impl::estimateClusterRadius (impl/CreateClustering.hpp:223-278) for an arbitrary bounding box."""
    a = as_xyz(xyz)
    lo = np.zeros(3)
    hi = np.zeros(3)
    lo[:len(bb_min)] = bb_min
    hi[:len(bb_max)] = bb_max
    r = C.c_double(0)
    _check(lib().orc_estimate_cluster_radius(_dp(a), C.c_int(a.shape[0]), C.c_int(dim), C.c_uint(vpc), _dp(lo), _dp(hi), C.byref(r)))
    return r.value


class PumMapping:
    """Mirror of mapping::PartitionOfUnityMapping (PartitionOfUnityMapping.hpp:48-57) on the oracle."""

    def __init__(self, constraint, dim, basis, param, polynomial="separate", vertices_per_cluster=50,
                 relative_overlap=0.15, project_to_input=True, gauss_support=None):
        self._l = lib()
        self.constraint, self.dim = constraint, dim
        self._h = self._l.orc_pum_create(C.c_int(CONSTRAINT[constraint]), C.c_int(dim), C.c_int(BASIS[basis]),
                                         C.c_double(param), C.c_double(_gauss_support(gauss_support)),
                                         C.c_int(POLYNOMIAL[polynomial]), C.c_uint(vertices_per_cluster),
                                         C.c_double(relative_overlap), C.c_int(int(project_to_input)))
        if not self._h:
            raise OracleError(self._l.orc_last_error().decode())
        self._h = C.c_void_p(self._h)
        self._n_in = self._n_out = 0

    def compute_mapping(self, input_xyz, output_xyz, threads=1):
        a, b = as_xyz(input_xyz), as_xyz(output_xyz)
        self._n_in, self._n_out = a.shape[0], b.shape[0]
        _check(self._l.orc_pum_compute(self._h, _dp(a), C.c_int(a.shape[0]), _dp(b), C.c_int(b.shape[0]), C.c_int(threads)))

    def compute_mapping_for_samples(self, input_xyz, output_xyz, samples, threads=1):
        """Full-size fixtures: builds only the clusters that hold one of the `samples` (vertex IDs of output_xyz, the mesh
        the result lives on); map() then returns the complete mapping's values at those vertices."""
        a, b = as_xyz(input_xyz), as_xyz(output_xyz)
        self._n_in, self._n_out = a.shape[0], b.shape[0]
        s = np.ascontiguousarray(samples, dtype=np.int32)
        _check(self._l.orc_pum_compute_sampled(self._h, _dp(a), C.c_int(a.shape[0]), _dp(b), C.c_int(b.shape[0]), _ip(s),
                                               C.c_int(s.size), C.c_int(threads)))

    def map(self, values, dims=1, threads=1) -> np.ndarray:
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
        assert v.size == self._n_in * dims, (v.size, self._n_in, dims)
        out = np.zeros(self._n_out * dims)
        _check(self._l.orc_pum_map(self._h, _dp(v), C.c_int(dims), _dp(out), C.c_int(threads)))
        return out

    # ---- just-in-time mapping (PartitionOfUnityMapping.hpp:382-476) ----
    def compute_mapping_jit(self, mesh_xyz, threads=1):
        """The other mesh is a just-in-time mesh; mesh_xyz is the mesh that exists (input for consistent, output for conservative)."""
        a = as_xyz(mesh_xyz)
        self._n_in, self._n_out = (a.shape[0], 0) if self.constraint != "conservative" else (0, a.shape[0])
        self._n_mesh = a.shape[0]
        _check(self._l.orc_pum_compute_jit(self._h, _dp(a), C.c_int(a.shape[0]), C.c_int(threads)))

    def initialize_cache(self, dims):
        c = self._l.orc_pum_cache_create(self._h, C.c_int(dims))
        if not c:
            raise OracleError(self._l.orc_last_error().decode())
        return {"h": C.c_void_p(c), "dims": dims}

    def reset_cache(self, cache):
        _check(self._l.orc_pum_cache_reset(self._h, cache["h"]))

    def update_cache(self, cache, values):
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
        assert v.size == self._n_mesh * cache["dims"]
        _check(self._l.orc_pum_cache_update(self._h, cache["h"], _dp(v)))

    def map_consistent_at(self, coords, cache) -> np.ndarray:
        x = as_xyz(coords)
        out = np.zeros(x.shape[0] * cache["dims"])
        _check(self._l.orc_pum_map_consistent_at(self._h, cache["h"], _dp(x), C.c_int(x.shape[0]), _dp(out)))
        return out

    def map_conservative_at(self, coords, source, cache):
        x = as_xyz(coords)
        v = np.ascontiguousarray(np.asarray(source, dtype=np.float64).reshape(-1))
        assert v.size == x.shape[0] * cache["dims"]
        _check(self._l.orc_pum_map_conservative_at(self._h, cache["h"], _dp(x), C.c_int(x.shape[0]), _dp(v)))

    def complete_just_in_time_mapping(self, cache, out):
        assert out.dtype == np.float64 and out.size == self._n_mesh * cache["dims"]
        _check(self._l.orc_pum_complete_jit(self._h, cache["h"], _dp(out)))

    def clear(self):
        self._l.orc_pum_clear(self._h)

    @property
    def cluster_radius(self):
        return self._l.orc_pum_radius(self._h)

    @property
    def num_clusters(self):
        return self._l.orc_pum_num_clusters(self._h)

    def cluster(self, c, with_cholesky=False):
        center = np.zeros(3)
        n_in, n_out, pp = C.c_int(0), C.c_int(0), C.c_int(0)
        _check(self._l.orc_pum_cluster_sizes(self._h, C.c_int(c), _dp(center), C.byref(n_in), C.byref(n_out), C.byref(pp)))
        in_ids = np.zeros(n_in.value, dtype=np.int32)
        out_ids = np.zeros(n_out.value, dtype=np.int32)
        w = np.zeros(n_out.value)
        L = np.zeros((n_in.value, n_in.value), order="F") if with_cholesky else None
        _check(self._l.orc_pum_cluster_data(self._h, C.c_int(c), _ip(in_ids), _ip(out_ids), _dp(w),
                                            _dp(L) if with_cholesky else None))
        d = {"center": center, "in_ids": in_ids, "out_ids": out_ids, "weights": w, "poly_params": pp.value}
        if with_cholesky:
            d["L"] = np.tril(L)
        return d

    def __del__(self):
        try:
            self._l.orc_pum_destroy(self._h)
        except Exception:
            pass


class GlobalMapping:
    """Mirror of mapping::RadialBasisFctMapping<RadialBasisFctSolver<RBF>> (RadialBasisFctMapping.hpp:40-46)."""

    def __init__(self, constraint, dim, basis, param=0.0, dead_axis=(False, False, False), polynomial="separate",
                 gauss_support=None):
        self._l = lib()
        dead = np.asarray([int(bool(x)) for x in dead_axis], dtype=np.int32)
        self._h = self._l.orc_global_create(C.c_int(CONSTRAINT[constraint]), C.c_int(dim), C.c_int(BASIS[basis]),
                                            C.c_double(param), C.c_double(_gauss_support(gauss_support)), _ip(dead),
                                            C.c_int(POLYNOMIAL[polynomial]))
        if not self._h:
            raise OracleError(self._l.orc_last_error().decode())
        self._h = C.c_void_p(self._h)
        self._n_in = self._n_out = 0

    def compute_mapping(self, input_xyz, output_xyz):
        a, b = as_xyz(input_xyz), as_xyz(output_xyz)
        self._n_in, self._n_out = a.shape[0], b.shape[0]
        _check(self._l.orc_global_compute(self._h, _dp(a), C.c_int(a.shape[0]), _dp(b), C.c_int(b.shape[0])))

    def map(self, values, dims=1) -> np.ndarray:
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
        assert v.size == self._n_in * dims
        out = np.zeros(self._n_out * dims)
        _check(self._l.orc_global_map(self._h, _dp(v), C.c_int(dims), _dp(out)))
        return out

    def clear(self):
        self._l.orc_global_clear(self._h)

    def __del__(self):
        try:
            self._l.orc_global_destroy(self._h)
        except Exception:
            pass


def cholesky(a: np.ndarray):
    m = np.asfortranarray(np.array(a, dtype=np.float64))
    rc = lib().orc_cholesky(_dp(m), C.c_int(m.shape[0]))
    return np.tril(m), rc == 0


def qr_solve(a: np.ndarray, b: np.ndarray, transposed=False):
    m = np.asfortranarray(np.array(a, dtype=np.float64))
    rhs = np.asfortranarray(np.array(b, dtype=np.float64).reshape(b.shape[0], -1))
    x = np.zeros((m.shape[0] if transposed else m.shape[1], rhs.shape[1]), order="F")
    rank = C.c_int(0)
    _check(lib().orc_qr_solve(_dp(m), C.c_int(m.shape[0]), C.c_int(m.shape[1]), _dp(rhs), C.c_int(rhs.shape[1]), _dp(x),
                              C.c_int(int(transposed)), C.byref(rank)))
    return x, rank.value


def singular_values(a: np.ndarray) -> np.ndarray:
    m = np.asfortranarray(np.array(a, dtype=np.float64))
    sv = np.zeros(min(m.shape))
    _check(lib().orc_singular_values(_dp(m), C.c_int(m.shape[0]), C.c_int(m.shape[1]), _dp(sv)))
    return sv


def pum_time(constraint, dim, basis, param, polynomial, vpc, overlap, project, in_xyz, out_xyz, values, dims, threads):
    """One computeMapping + one map on `threads` host cores; returns (t_compute, t_map, out, n_clusters)."""
    a, b = as_xyz(in_xyz), as_xyz(out_xyz)
    v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).reshape(-1))
    n_res = (a.shape[0] if constraint == "conservative" else b.shape[0]) if False else b.shape[0]
    out = np.zeros(n_res * dims)
    t = np.zeros(2)
    ncl = C.c_int(0)
    _check(lib().orc_pum_time(C.c_int(CONSTRAINT[constraint]), C.c_int(dim), C.c_int(BASIS[basis]), C.c_double(param),
                              C.c_int(POLYNOMIAL[polynomial]), C.c_uint(vpc), C.c_double(overlap), C.c_int(int(project)),
                              _dp(a), C.c_int(a.shape[0]), _dp(b), C.c_int(b.shape[0]), _dp(v), C.c_int(dims), _dp(out),
                              C.c_int(threads), _dp(t), C.byref(ncl)))
    return t[0], t[1], out, ncl.value


def pum_time_ranks(constraint, dim, basis, param, polynomial, vpc, overlap, project, in_parts, out_parts, val_parts, dims):
    """One independent computeMapping + map per partition, each on its own host thread (one preCICE rank per core).
    Returns (t_compute, t_map, [out per partition], n_clusters)."""
    ins = [as_xyz(a) for a in in_parts]
    outs = [as_xyz(a) for a in out_parts]
    a, b = np.ascontiguousarray(np.concatenate(ins)), np.ascontiguousarray(np.concatenate(outs))
    v = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in val_parts]))
    in_off = np.zeros(len(ins) + 1, dtype=np.int64)
    out_off = np.zeros(len(outs) + 1, dtype=np.int64)
    in_off[1:] = np.cumsum([x.shape[0] for x in ins])
    out_off[1:] = np.cumsum([x.shape[0] for x in outs])
    out = np.zeros(b.shape[0] * dims)
    t = np.zeros(2)
    ncl = C.c_longlong(0)
    _check(lib().orc_pum_time_ranks(C.c_int(CONSTRAINT[constraint]), C.c_int(dim), C.c_int(BASIS[basis]), C.c_double(param),
                                    C.c_int(POLYNOMIAL[polynomial]), C.c_uint(vpc), C.c_double(overlap), C.c_int(int(project)),
                                    _dp(a), in_off.ctypes.data_as(C.POINTER(C.c_longlong)), _dp(b),
                                    out_off.ctypes.data_as(C.POINTER(C.c_longlong)), _dp(v), C.c_int(dims), _dp(out),
                                    C.c_int(len(ins)), _dp(t), C.byref(ncl)))
    res = [out[out_off[r] * dims:out_off[r + 1] * dims] for r in range(len(ins))]
    return t[0], t[1], res, ncl.value


__all__ = ["pum_time_ranks", "BASIS", "POLYNOMIAL", "CONSTRAINT", "OracleError", "build", "lib", "as_xyz", "basis_eval", "basis_params",
           "create_clustering", "PumMapping", "GlobalMapping", "cholesky", "qr_solve", "singular_values", "pum_time"]
_ = math
