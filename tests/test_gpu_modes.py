"""`-m gpu`: execution-mode minimal-compute (computeEvaluationOffline = true, PartitionOfUnityMapping.hpp:48-57;
MappingConfiguration.cpp:328-331, 555-559; BatchedRBFSolver.hpp:196-239, 324) against minimal-memory and the CPU oracle:
the stored evaluation matrices hold exactly the values the minimal-memory kernels re-evaluate, so the two modes agree to
summation order (1e-13), and both agree with the oracle at 1e-10 (north_star)."""
import numpy as np
import pytest

import refcases

pytestmark = pytest.mark.gpu
REL_L2_TOL = 1e-10


@pytest.fixture(scope="module")
def pb():
    import precice_b200
    if precice_b200.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need a B200 (no CPU fallback exists)")
    return precice_b200


def _case(dim, n_in, n_out, dims):
    inp, out = refcases.halton(n_in, (2, 3, 5)), refcases.halton(n_out, (7, 11, 13))
    if dim == 2:
        inp[:, 2] = 0.0
        out[:, 2] = 0.0
    f = np.sin(2 * np.pi * inp[:, 0]) * np.cos(2 * np.pi * inp[:, 1]) + inp[:, 2]
    vals = f if dims == 1 else np.stack([f * (k + 1) + k for k in range(dims)], axis=1).reshape(-1)
    return inp, out, vals


@pytest.mark.parametrize("constraint", ["consistent", "conservative"])
@pytest.mark.parametrize("dim,basis,n_in,n_out,dims,vpc", [
    (3, "compact-polynomial-c6", 5000, 4200, 1, 50),
    (3, "compact-polynomial-c2", 4000, 5500, 3, 30),   # more output than input vertices per cluster
    (2, "compact-polynomial-c4", 3000, 3000, 2, 40),
    (3, "gaussian", 3000, 2600, 4, 25),               # run-time dispatched basis function, 4 components = two launches
    (3, "compact-polynomial-c6", 2500, 2500, 1, 90),  # clusters of 64 < n <= 128 vertices (tiled kernels)
    (3, "compact-polynomial-c0", 1500, 1500, 2, 200),  # clusters of more than 128 vertices (global-memory kernels)
])
def test_minimal_compute_matches_minimal_memory_and_oracle(pb, constraint, dim, basis, n_in, n_out, dims, vpc):
    import oracle
    inp, out, vals = _case(dim, n_in, n_out, dims)
    probe = oracle.PumMapping(constraint, dim, "compact-polynomial-c6", 1.0, "separate", vpc, 0.15, True)
    probe.compute_mapping(inp, out)
    param = 2.0 * probe.cluster_radius
    if basis == "gaussian":
        param = 1.2 / probe.cluster_radius
    args = (constraint, dim, basis, param, "separate", vpc, 0.15, True)
    mm = pb.PartitionOfUnityMapping(*args, execution_mode="minimal-memory")
    mc = pb.PartitionOfUnityMapping(*args, execution_mode="minimal-compute")
    mm.computeMapping(inp, out)
    mc.computeMapping(inp, out)
    st_mm, st_mc = mm.stats(), mc.stats()
    assert st_mm["evaluation_bytes"] == 0
    # lane side rounded up to even, times the loop side, per cluster
    _, _, in_off, out_off, _, _ = mc.clusters()
    n_c, m_c = np.diff(in_off), np.diff(out_off)
    lane, loop = (m_c, n_c) if constraint == "consistent" else (n_c, m_c)
    assert st_mc["evaluation_bytes"] == 8 * int((((lane + 1) // 2 * 2) * loop).sum())
    r_mm, r_mc = mm.map(vals, dims), mc.map(vals, dims)
    assert refcases.rel_l2(r_mc, r_mm) <= 1e-13
    assert refcases.rel_l2(mc.map(vals, dims), r_mc) <= 1e-15  # a second map reuses the stored matrices
    mc.clear()
    mc.computeMapping(inp, out)  # clear + recompute rebuilds them
    assert refcases.rel_l2(mc.map(vals, dims), r_mm) <= 1e-13
    orc = oracle.PumMapping(*args)
    orc.compute_mapping(inp, out)
    tol = REL_L2_TOL if basis != "gaussian" else 1e-8  # cond(C_c) ~ 1e6 for this gaussian; both modes share the factor
    assert refcases.rel_l2(r_mc, orc.map(vals, dims)) <= tol


def test_minimal_compute_sharded_virtual_ranks(pb):
    n, vpc = 6000, 30
    inp, out, vals = _case(3, n, n, 1)
    support = 2.0 * (3.0 * vpc / (4.0 * np.pi * n)) ** (1.0 / 3.0)
    args = ("consistent", 3, "compact-polynomial-c6", support, "separate", vpc, 0.15, True)
    single = pb.PartitionOfUnityMapping(*args)
    single.computeMapping(inp, out)
    ref = single.map(vals, 1)
    total = np.zeros_like(ref)
    for r in range(3):
        m = pb.PartitionOfUnityMapping(*args, execution_mode="minimal-compute")
        m.distributedInit(None, r, 3)
        m.computeMapping(inp, out)
        total += m.map(vals, 1)
    assert refcases.rel_l2(total, ref) <= 1e-13


def test_unknown_modes_are_rejected(pb):
    from precice_b200 import _lib
    import ctypes as C
    cfg = _lib.Config()
    L = _lib.lib()
    L.rbfmap_default_config(C.byref(cfg))
    cfg.support_radius = 0.5
    for field in ("execution_mode", "shard_mode"):
        setattr(cfg, field, 7)
        h = C.c_void_p()
        assert L.rbfmap_create(C.byref(cfg), C.byref(h)) == 1  # RBFMAP_ERR_INVALID
        setattr(cfg, field, 0)
