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
    assert refcases.rel_l2(r_mc, r_mm) <= 1e-12  # same entries, different summation order (times cond(C_c) in the conservative direction)
    assert refcases.rel_l2(mc.map(vals, dims), r_mc) <= 1e-13  # a second map reuses the stored matrices
    mc.clear()
    mc.computeMapping(inp, out)  # clear + recompute rebuilds them
    assert refcases.rel_l2(mc.map(vals, dims), r_mm) <= 1e-12
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


# ---------------------------------------------------------------------------------------------- scaled-consistent
def _integral(xyz, elements, values, dims, dim):
    """mesh::integrateSurface / integrateVolume (mesh/Utils.cpp:11-70) in the reference's element order."""
    total = np.zeros(dims)
    v = np.asarray(values).reshape(-1, dims)
    for e in elements:
        p = xyz[list(e), :dim]
        if len(e) == 2:
            measure = np.linalg.norm(p[1] - p[0])
        elif len(e) == 3:
            a, b = p[1] - p[0], p[2] - p[0]
            measure = 0.5 * abs(a[0] * b[1] - a[1] * b[0]) if dim == 2 else 0.5 * np.linalg.norm(np.cross(a, b))
        else:
            measure = abs(np.dot(p[0] - p[3], np.cross(p[1] - p[3], p[2] - p[3]))) / 6.0
        total += measure / len(e) * v[list(e)].sum(0)
    return total


SCALED_MAPPINGS = [
    lambda pb, c, d: pb.RadialBasisFctMapping(c, d, "thin-plate-splines", polynomial="separate"),
    lambda pb, c, d: pb.RadialBasisFctMapping(c, d, "compact-polynomial-c6", 5.0, polynomial="separate"),
    lambda pb, c, d: pb.RadialBasisFctMapping(c, d, "gaussian", 1.0, polynomial="separate", solver="iterative"),
    lambda pb, c, d: pb.PartitionOfUnityMapping(c, d, "compact-polynomial-c6", 5.0, "separate", 50, 0.15, True),
]


@pytest.mark.parametrize("make", SCALED_MAPPINGS)
def test_scaled_consistent_2d(pb, make):
    """perform2DTestScaledConsistentMapping (mapping/tests/RadialBasisFctHelper.hpp:397-447): edges, scalar data."""
    inp = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], float)
    out = np.array([[0, 0, 0], [0, 1, 0], [1.1, 1.1, 0], [0.1, 1.1, 0]], float)
    e_in, e_out = [(0, 1), (1, 2), (2, 3), (0, 3)], [(0, 1), (1, 2), (2, 3), (0, 3)]
    vals = np.array([1.0, 2.0, 2.0, 1.0])
    m = make(pb, "scaled-consistent-surface", 2)
    m.setConnectivity("input", e_in)
    m.setConnectivity("output", e_out)
    m.computeMapping(inp, out)
    res = m.map(vals, 1)
    assert _integral(out, e_out, res, 1, 2) == pytest.approx(_integral(inp, e_in, vals, 1, 2), rel=1e-12)
    # the scaling is one factor per component on top of the consistent result
    plain = make(pb, "consistent", 2)
    plain.computeMapping(inp, out)
    ratio = res / plain.map(vals, 1)
    assert np.allclose(ratio, ratio[0], rtol=1e-12) and abs(ratio[0] - 1.0) > 1e-3


@pytest.mark.parametrize("make", SCALED_MAPPINGS)
def test_scaled_consistent_3d_surface_vector(pb, make):
    """perform3DTestScaledConsistentMapping (:449-505), with a 2-component field: one factor per component."""
    inp = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.5], [2, 0, 0], [0, 2, 0], [0, 2, 1]], float)
    out = np.array([[0, 0, 0], [1, 0, 0], [0, 1.1, 0.6]], float)
    t_in, t_out = [(0, 1, 2), (3, 4, 5)], [(0, 1, 2)]
    vals = np.stack([[1.0, 2.0, 4.0, 6.0, 8.0, 9.0], [3.0, -1.0, 0.5, 2.0, 2.0, 7.0]], axis=1).reshape(-1)
    m = make(pb, "scaled-consistent-surface", 3)
    m.setConnectivity("input", t_in)
    m.setConnectivity("output", t_out)
    m.computeMapping(inp, out)
    res = m.map(vals, 2)
    assert _integral(out, t_out, res, 2, 3) == pytest.approx(_integral(inp, t_in, vals, 2, 3), rel=1e-12)


def test_scaled_consistent_volume_and_errors(pb):
    rng = np.random.default_rng(11)
    inp, out = rng.random((40, 3)), rng.random((30, 3))
    tets_in = [tuple(rng.choice(40, 4, replace=False)) for _ in range(25)]
    tets_out = [tuple(rng.choice(30, 4, replace=False)) for _ in range(20)]
    vals = rng.random(40) + 1.0
    m = pb.RadialBasisFctMapping("scaled-consistent-volume", 3, "thin-plate-splines", polynomial="separate")
    m.setConnectivity("input", tets_in)
    m.computeMapping(inp, out)
    with pytest.raises(pb.MappingError) as e:  # Mapping.cpp:204-206
        m.map(vals, 1)
    assert 'Tetrahedra connectivity information is missing for the mesh "output"' in str(e.value)
    with pytest.raises(pb.MappingError) as e:  # the weights belong to the computed mapping
        m.setConnectivity("output", tets_out)
    assert e.value.status == "STATE"
    m.clear()
    m.setConnectivity("output", tets_out)
    m.computeMapping(inp, out)
    res = m.map(vals, 1)
    assert _integral(out, tets_out, res, 1, 3) == pytest.approx(_integral(inp, tets_in, vals, 1, 3), rel=1e-12)
    # triangles where tetrahedra are required count as missing
    m2 = pb.RadialBasisFctMapping("scaled-consistent-volume", 3, "thin-plate-splines", polynomial="separate")
    m2.setConnectivity("input", [(0, 1, 2)])
    m2.setConnectivity("output", [(0, 1, 2)])
    m2.computeMapping(inp, out)
    with pytest.raises(pb.MappingError) as e:
        m2.map(vals, 1)
    assert "Tetrahedra connectivity information is missing" in str(e.value)


# ---------------------------------------------------------------------------------------------- non-SPD basis functions in PUM
@pytest.mark.parametrize("constraint", ["consistent", "conservative"])
@pytest.mark.parametrize("dim,basis,param,polynomial,dims,vpc", [
    (3, "thin-plate-splines", 0.0, "separate", 1, 40),
    (3, "multiquadrics", 0.05, "separate", 3, 30),
    (3, "volume-splines", 0.0, "off", 1, 50),
    (2, "thin-plate-splines", 0.0, "separate", 2, 35),
    (3, "volume-splines", 0.0, "separate", 1, 90),   # clusters of 64 < n <= 128 vertices
])
def test_pum_with_functions_that_are_not_positive_definite(pb, constraint, dim, basis, param, polynomial, dims, vpc):
    """RadialBasisFctSolver.hpp:354-358: thin-plate splines, multiquadrics and volume splines take the ColPivHouseholderQR
    branch in every cluster (SphericalVertexCluster.hpp:139-182); the device builds the explicit inverse instead
    (pum_inverse.cu). Tolerance 1e-7: the one the global variants use for the same functions (cond(C_c) up to ~1e8;
    test_extended_precision_bounds_the_tolerances shows the oracle itself is no closer to the exact solution)."""
    import oracle
    n_in, n_out = 3500, 3100
    inp, out, vals = _case(dim, n_in, n_out, dims)
    args = (constraint, dim, basis, param, polynomial, vpc, 0.15, True)
    for mode in ("minimal-memory", "minimal-compute"):
        m = pb.PartitionOfUnityMapping(*args, execution_mode=mode)
        m.computeMapping(inp, out)
        res = m.map(vals, dims)
        if mode == "minimal-memory":
            orc = oracle.PumMapping(*args)
            orc.compute_mapping(inp, out)
            ref = orc.map(vals, dims)
            assert m.stats()["n_clusters"] == orc.num_clusters
        assert refcases.rel_l2(res, ref) <= 1e-7, (mode, refcases.rel_l2(res, ref))
        if constraint == "conservative" and polynomial == "separate":
            assert np.allclose(res.reshape(-1, dims).sum(0), vals.reshape(-1, dims).sum(0), rtol=1e-9)


def test_pum_not_positive_definite_limits(pb):
    inp, out, vals = _case(3, 1200, 1200, 1)
    m = pb.PartitionOfUnityMapping("consistent", 3, "thin-plate-splines", 0.0, "separate", 300, 0.15, True)
    with pytest.raises(pb.MappingError) as e:
        m.computeMapping(inp, out)
    assert e.value.status == "UNSUPPORTED" and "128 vertices" in str(e.value)
    for r in range(2):  # a rank of a sharded mapping reaches the ranks' agreement first and raises the same error there
        m = pb.PartitionOfUnityMapping("consistent", 3, "thin-plate-splines", 0.0, "separate", 300, 0.15, True)
        m.distributedInit(None, r, 2)
        with pytest.raises(pb.MappingError) as e:
            m.computeMapping(inp, out)
        assert e.value.status == "UNSUPPORTED" and "128 vertices" in str(e.value)
    # duplicated vertices make the cluster matrices singular: the reference's message (RadialBasisFctSolver.hpp:360-365)
    dup = np.concatenate([inp, inp[:40]])
    m = pb.PartitionOfUnityMapping("consistent", 3, "volume-splines", 0.0, "separate", 40, 0.15, True)
    with pytest.raises(pb.MappingError) as e:
        m.computeMapping(dup, out)
    assert e.value.status == "SINGULAR" and "is not invertable" in str(e.value)
    with pytest.raises(pb.MappingError) as e:
        pb.PartitionOfUnityMapping("consistent", 3, "multiquadrics", 0.1).computeMappingJustInTime(inp)
    assert e.value.status == "UNSUPPORTED"


# ---------------------------------------------------------------------------------------------- deterministic blend
@pytest.mark.parametrize("constraint", ["consistent", "conservative"])
@pytest.mark.parametrize("execution_mode", ["minimal-memory", "minimal-compute"])
def test_deterministic_blend_is_bit_reproducible(pb, constraint, execution_mode):
    """rbfmap_config.deterministic: weight sums and blend are gathered per vertex in ascending cluster order (the order of
    the reference's loop over the clusters, SphericalVertexCluster.hpp:237, 332) instead of atomics: two independent
    mappings give bit-identical results, which the atomic version does not guarantee."""
    import oracle
    inp, out, vals = _case(3, 9000, 8000, 3)
    args = (constraint, 3, "compact-polynomial-c6", 2.0 * (3.0 * 40 / (4.0 * np.pi * 9000)) ** (1.0 / 3.0), "separate", 40, 0.15, True)
    runs = []
    for _ in range(3):
        m = pb.PartitionOfUnityMapping(*args, deterministic=True, execution_mode=execution_mode)
        m.computeMapping(inp, out)
        runs.append((m.map(vals, 3), m.map(vals[::3].copy(), 1)))
        del m
    for a, b in runs[1:]:
        assert np.array_equal(a, runs[0][0]) and np.array_equal(b, runs[0][1])
    plain = pb.PartitionOfUnityMapping(*args)
    plain.computeMapping(inp, out)
    assert refcases.rel_l2(runs[0][0], plain.map(vals, 3)) <= 1e-12
    orc = oracle.PumMapping(*args)
    orc.compute_mapping(inp, out)
    assert refcases.rel_l2(runs[0][0], orc.map(vals, 3)) <= REL_L2_TOL


def test_deterministic_blend_sharded(pb):
    inp, out, vals = _case(3, 6000, 6000, 1)
    args = ("consistent", 3, "compact-polynomial-c6", 2.0 * (3.0 * 30 / (4.0 * np.pi * 6000)) ** (1.0 / 3.0), "separate", 30, 0.15, True)
    single = pb.PartitionOfUnityMapping(*args, deterministic=True)
    single.computeMapping(inp, out)
    ref = single.map(vals, 1)
    total = np.zeros_like(ref)
    for r in range(3):
        m = pb.PartitionOfUnityMapping(*args, deterministic=True)
        m.distributedInit(None, r, 3)
        m.computeMapping(inp, out)
        total += m.map(vals, 1)
    # every vertex is owned by one rank, which sums the same clusters in the same order; inside a boundary cluster the
    # evaluation loop splits its lanes by the number of (owned) output vertices, so the last bits may differ
    assert refcases.rel_l2(total, ref) <= 1e-15
    again = np.zeros_like(ref)
    for r in range(3):
        m = pb.PartitionOfUnityMapping(*args, deterministic=True)
        m.distributedInit(None, r, 3)
        m.computeMapping(inp, out)
        again += m.map(vals, 1)
    assert np.array_equal(again, total)  # ... but the sharded mapping itself is bit-reproducible


# ---------------------------------------------------------------------------------------------- batched samples
@pytest.mark.parametrize("constraint", ["consistent", "conservative"])
def test_map_batch_equals_separate_maps(pb, constraint):
    """rbfmap_map_batch: the stamples / data of DataContext::mapData (DataContext.cpp:110-171) in one call."""
    inp, out, f = _case(3, 5000, 4500, 1)
    samples = [f, np.stack([f, 2 * f + 1, -f], 1).reshape(-1), np.cos(5 * f), np.stack([f * f, 1 - f], 1).reshape(-1)]
    dims = [1, 3, 1, 2]
    makers = [lambda: pb.PartitionOfUnityMapping(constraint, 3, "compact-polynomial-c6", 0.2, "separate", 50, 0.15, True),
              lambda: pb.RadialBasisFctMapping(constraint, 3, "compact-polynomial-c2", 0.3, polynomial="separate")]
    for make in makers:
        m = make()
        m.computeMapping(inp, out)
        separate = [m.map(s, d) for s, d in zip(samples, dims)]
        batched = m.mapBatch(samples, dims)
        for a, b in zip(batched, separate):
            assert a.shape == b.shape and refcases.rel_l2(a, b) <= 1e-13
        assert m.mapBatch([], []) == []
