"""ctypes binding of the C ABI in include/rbfmap.h (precice_b200/librbfmap.so).

The shared library is the product; this module only loads it and declares the signatures. There is
no CPU fallback: if the library is missing, loading fails loudly; if there is no CUDA device, every
compute call returns RBFMAP_ERR_CUDA and raises MappingError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librbfmap.so")

# enums of include/rbfmap.h
POLYNOMIAL = {"on": 0, "off": 1, "separate": 2}
BASIS = {
    "compact-polynomial-c0": 0, "compact-polynomial-c2": 1, "compact-polynomial-c4": 2,
    "compact-polynomial-c6": 3, "compact-polynomial-c8": 4, "thin-plate-splines": 5,
    "multiquadrics": 6, "inverse-multiquadrics": 7, "volume-splines": 8, "gaussian": 9,
    "compact-tps-c2": 10,
}
CONSTRAINT = {"consistent": 0, "conservative": 1, "scaled-consistent-surface": 2, "scaled-consistent-volume": 3}
VARIANT = {"rbf-global-direct": 0, "rbf-global-iterative": 1, "rbf-pum-direct": 2}
EXECUTION_MODE = {"minimal-memory": 0, "minimal-compute": 1}
SHARD_MODE = {"duplicate": 0, "exchange": 1}
STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "SINGULAR", 4: "UNASSIGNED", 5: "EMPTY", 6: "STATE", 7: "UNSUPPORTED",
          8: "NOT_CONVERGED"}

# every symbol include/rbfmap.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "rbfmap_default_config", "rbfmap_create", "rbfmap_compute_mapping", "rbfmap_map", "rbfmap_map_consistent",
    "rbfmap_map_conservative", "rbfmap_compute_mapping_device", "rbfmap_map_device", "rbfmap_clear",
    "rbfmap_has_computed_mapping", "rbfmap_destroy", "rbfmap_get_name", "rbfmap_last_error", "rbfmap_get_stats",
    "rbfmap_pum_get_clusters", "rbfmap_pum_get_cluster_ids", "rbfmap_eval_basis", "rbfmap_device_count",
    "rbfmap_device_info", "rbfmap_measure_fp64_peak", "rbfmap_debug_dist_sqrt", "rbfmap_debug_pair_basis", "rbfmap_nccl_unique_id", "rbfmap_distributed_init",
    "rbfmap_partition_rows", "rbfmap_partition_slabs", "rbfmap_struct_sizes", "rbfmap_tag_mesh_first_round", "rbfmap_tag_mesh_second_round",
    "rbfmap_estimate_cluster_radius", "rbfmap_set_connectivity", "rbfmap_map_batch", "rbfmap_map_sliced", "rbfmap_initial_guess_size", "rbfmap_map_with_initial_guess", "rbfmap_host_alloc", "rbfmap_host_free", "rbfmap_compute_mapping_just_in_time", "rbfmap_cache_create", "rbfmap_cache_reset", "rbfmap_cache_destroy",
    "rbfmap_cache_update", "rbfmap_map_consistent_at", "rbfmap_map_conservative_at", "rbfmap_complete_just_in_time_mapping",
]


class Config(C.Structure):
    _fields_ = [
        ("variant", C.c_int), ("constraint", C.c_int), ("dimensions", C.c_int), ("basis", C.c_int),
        ("support_radius", C.c_double), ("shape_parameter", C.c_double), ("polynomial", C.c_int),
        ("dead_axis", C.c_int * 3), ("vertices_per_cluster", C.c_uint), ("relative_overlap", C.c_double),
        ("project_to_input", C.c_int), ("solver_rtol", C.c_double), ("max_iterations", C.c_int), ("device_id", C.c_int),
        ("execution_mode", C.c_int), ("shard_mode", C.c_int), ("rescaling", C.c_int), ("output_filter", C.c_double),
        ("deterministic", C.c_int),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_in", C.c_int64), ("n_out", C.c_int64), ("n_clusters", C.c_int64), ("cluster_radius", C.c_double),
        ("sum_n", C.c_int64), ("sum_m", C.c_int64), ("sum_n2", C.c_int64), ("sum_n3", C.c_int64), ("sum_mn", C.c_int64),
        ("max_n", C.c_int64), ("device_bytes", C.c_int64), ("iterations", C.c_int), ("residual", C.c_double),
        ("ms_compute_kernels", C.c_float), ("ms_map_kernels", C.c_float), ("ms_assembly_factor", C.c_float),
        ("kernel_launches", C.c_int),
        ("n_clusters_global", C.c_int64), ("shard_axis", C.c_int), ("shard_lo", C.c_double), ("shard_hi", C.c_double),
        ("ms_replicated", C.c_float), ("ms_collectives", C.c_float), ("evaluation_bytes", C.c_int64),
        ("ms_evaluation_offline", C.c_float),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_lib = None


def build():
    """Compiles librbfmap.so in-tree with nvcc for sm_100a (used by __graft_entry__.build())."""
    import subprocess
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). precice_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        vp = C.c_void_p
        L.rbfmap_default_config.argtypes = [C.POINTER(Config)]
        L.rbfmap_default_config.restype = None
        L.rbfmap_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.rbfmap_compute_mapping.argtypes = [vp, vp, C.c_int64, vp, C.c_int64]
        L.rbfmap_compute_mapping_device.argtypes = [vp, vp, C.c_int64, vp, C.c_int64]
        for f in (L.rbfmap_map, L.rbfmap_map_consistent, L.rbfmap_map_conservative, L.rbfmap_map_device):
            f.argtypes = [vp, vp, C.c_int, vp]
        L.rbfmap_clear.argtypes = [vp]
        L.rbfmap_has_computed_mapping.argtypes = [vp]
        L.rbfmap_destroy.argtypes = [vp]
        L.rbfmap_destroy.restype = None
        L.rbfmap_get_name.argtypes = [vp]
        L.rbfmap_get_name.restype = C.c_char_p
        L.rbfmap_last_error.argtypes = [vp]
        L.rbfmap_last_error.restype = C.c_char_p
        L.rbfmap_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.rbfmap_pum_get_clusters.argtypes = [vp, vp, vp, vp]
        L.rbfmap_pum_get_cluster_ids.argtypes = [vp, vp, vp]
        L.rbfmap_eval_basis.argtypes = [C.c_int, C.c_double, C.c_double, dp, C.c_int64, dp]
        L.rbfmap_device_count.argtypes = []
        L.rbfmap_device_info.argtypes = [C.c_int, ip, ip, ip, C.POINTER(C.c_int64)]
        L.rbfmap_measure_fp64_peak.argtypes = [dp, dp]
        L.rbfmap_debug_dist_sqrt.argtypes = [dp, C.c_int64, dp]
        L.rbfmap_debug_pair_basis.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, dp, dp, C.c_int64, C.c_int, dp]
        L.rbfmap_nccl_unique_id.argtypes = [C.c_char_p]
        L.rbfmap_distributed_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
        L.rbfmap_partition_rows.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.rbfmap_partition_slabs.argtypes = [C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, ip]
        L.rbfmap_compute_mapping_just_in_time.argtypes = [vp, vp, C.c_int64]
        L.rbfmap_cache_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.rbfmap_cache_reset.argtypes = [vp, vp]
        L.rbfmap_cache_destroy.argtypes = [vp]
        L.rbfmap_cache_destroy.restype = None
        L.rbfmap_cache_update.argtypes = [vp, vp, vp]
        L.rbfmap_map_consistent_at.argtypes = [vp, vp, vp, C.c_int64, vp]
        L.rbfmap_map_conservative_at.argtypes = [vp, vp, vp, C.c_int64, vp]
        L.rbfmap_complete_just_in_time_mapping.argtypes = [vp, vp, vp]
        L.rbfmap_tag_mesh_first_round.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, vp]
        L.rbfmap_tag_mesh_second_round.argtypes = [vp, vp, C.c_int64, vp, vp]
        L.rbfmap_estimate_cluster_radius.argtypes = [vp, vp, C.c_int64, dp, dp, dp]
        L.rbfmap_set_connectivity.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int]
        L.rbfmap_map_sliced.argtypes = [vp, vp, C.c_int, vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.rbfmap_map_batch.argtypes = [vp, C.c_int, C.POINTER(vp), ip, C.POINTER(vp)]
        L.rbfmap_initial_guess_size.argtypes = [vp]
        L.rbfmap_initial_guess_size.restype = C.c_int64
        L.rbfmap_map_with_initial_guess.argtypes = [vp, vp, C.c_int, vp, vp]
        L.rbfmap_host_alloc.argtypes = [C.POINTER(vp), C.c_int64]
        L.rbfmap_host_free.argtypes = [vp]
        L.rbfmap_host_free.restype = None
        L.rbfmap_struct_sizes.argtypes = [ip, ip]
        L.rbfmap_struct_sizes.restype = None
        a, b = C.c_int(0), C.c_int(0)
        L.rbfmap_struct_sizes(C.byref(a), C.byref(b))
        if (a.value, b.value) != (C.sizeof(Config), C.sizeof(Stats)):
            raise ImportError(f"{LIB_PATH}: struct layout mismatch (library {a.value}/{b.value} bytes, binding "
                              f"{C.sizeof(Config)}/{C.sizeof(Stats)}); rebuild the library")
        _lib = L
    return _lib
