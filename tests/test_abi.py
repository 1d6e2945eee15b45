"""`-m "not gpu"`: the C-ABI library loads, exports every symbol include/rbfmap.h declares, validates configurations
on the host, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rbfmap.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rbfmap_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from precice_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared_symbols()


def test_library_exports_every_declared_symbol():
    from precice_b200 import _lib
    L = _lib.lib()
    for name in _declared_symbols():
        assert hasattr(L, name), name


def test_struct_layouts_match_header():
    from precice_b200 import _lib
    cfg = _lib.Config()
    _lib.lib().rbfmap_default_config(C.byref(cfg))
    # XML defaults: MappingConfiguration.cpp:227-248
    assert cfg.polynomial == _lib.POLYNOMIAL["separate"]
    assert cfg.vertices_per_cluster == 50
    assert cfg.relative_overlap == 0.15
    assert cfg.project_to_input == 1
    assert cfg.solver_rtol == 1e-9
    assert cfg.device_id == -1
    assert cfg.execution_mode == _lib.EXECUTION_MODE["minimal-memory"]  # PartitionOfUnityMapping.hpp:57
    assert cfg.shard_mode == _lib.SHARD_MODE["duplicate"] and cfg.deterministic == 0
    assert cfg.rescaling == 1 and cfg.output_filter == 1e-9  # PetRadialBasisFctMapping.hpp:126-127, 675
    assert C.sizeof(_lib.Config) == 120
    assert C.sizeof(_lib.Stats) == 176
    a, b = C.c_int(0), C.c_int(0)
    _lib.lib().rbfmap_struct_sizes(C.byref(a), C.byref(b))  # what the compiler laid out
    assert (a.value, b.value) == (C.sizeof(_lib.Config), C.sizeof(_lib.Stats))


def test_enum_values_follow_reference():
    from precice_b200 import _lib
    # mapping/config/MappingConfigurationTypes.hpp:11-30, mapping/Mapping.hpp:30-35
    assert _lib.POLYNOMIAL == {"on": 0, "off": 1, "separate": 2}
    assert _lib.BASIS["compact-polynomial-c0"] == 0 and _lib.BASIS["compact-polynomial-c8"] == 4
    assert _lib.BASIS["thin-plate-splines"] == 5 and _lib.BASIS["compact-tps-c2"] == 10
    assert _lib.CONSTRAINT == {"consistent": 0, "conservative": 1, "scaled-consistent-surface": 2, "scaled-consistent-volume": 3}


def test_no_cpu_fallback():
    import precice_b200 as pb
    if pb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pb.MappingError) as e:
        pb.PartitionOfUnityMapping("consistent", 3, "compact-polynomial-c6", 1.0)
    assert e.value.status == "CUDA"
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(pb.MappingError):
        pb.eval_basis("compact-polynomial-c6", 1.0, np.array([0.5]))


def test_constructor_argument_checks_run_on_host():
    """Invalid configurations are rejected with the reference's messages before any device is touched."""
    import precice_b200 as pb
    with pytest.raises(pb.MappingError) as e:  # BasisFunctions.hpp:519-523
        pb.PartitionOfUnityMapping("consistent", 3, "compact-polynomial-c6", 0.0)
    assert e.value.status == "INVALID" and "has to be larger than zero" in str(e.value)
    with pytest.raises(pb.MappingError) as e:  # BasisFunctions.hpp:233-236
        pb.PartitionOfUnityMapping("consistent", 3, "gaussian", -1.0)
    assert "Shape parameter for radial-basis-function gaussian" in str(e.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under precice_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "precice_b200")):
        if "_obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "rbf_oracle" not in text, os.path.join(dirpath, f)


def test_partition_slabs_host_helper():
    """Slab cuts of the sharded rbf-pum-direct: monotone, cover all bins, and no other choice of cuts has a
    smaller maximum load (brute force on small cases)."""
    import itertools
    from precice_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(5)

    def cuts_of(hist, world, halo):
        h = np.ascontiguousarray(hist, dtype=np.int64)
        c = (C.c_int * (world + 1))()
        assert L.rbfmap_partition_slabs(h.ctypes.data_as(C.POINTER(C.c_int64)), len(h), world, halo, c) == 0
        return list(c)

    def max_load(hist, cuts, halo):
        P = np.concatenate([[0], np.cumsum(hist)])
        n = len(hist)
        return max(P[min(n, b + halo)] - P[max(0, a - halo)] for a, b in zip(cuts, cuts[1:]))

    for n_bins, world, halo in ((12, 3, 0), (12, 3, 1), (10, 4, 2), (9, 2, 0), (8, 8, 0), (6, 1, 3)):
        for _ in range(5):
            hist = rng.integers(0, 20, n_bins)
            cuts = cuts_of(hist, world, halo)
            assert cuts[0] == 0 and cuts[-1] == n_bins and all(a <= b for a, b in zip(cuts, cuts[1:]))
            best = min(max_load(hist, [0, *inner, n_bins], halo)
                       for inner in itertools.combinations_with_replacement(range(n_bins + 1), world - 1))
            assert max_load(hist, cuts, halo) == best
    # uniform histogram: equal slabs
    cuts = cuts_of(np.full(4096, 230), 8, 0)
    assert cuts == [512 * r for r in range(9)]


def test_partition_rows_host_helper():
    """Row slices of the sharded CG: contiguous, equal length except the tail, cover [0, n)."""
    import precice_b200 as pb
    for n, world in ((10, 3), (5, 8), (500000, 8), (7, 1), (0, 4)):
        parts = [pb.partition_rows(n, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
            assert e0 == b1 and e0 - b0 >= e1 - b1
        assert max(e - b for b, e in parts) == -(-n // world)
