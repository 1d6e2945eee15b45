"""Pins the CPU oracle against the reference's own golden literals (`-m "not gpu"`)."""
import numpy as np
import pytest

import oracle
import refcases

CASES = refcases.pum_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_pum_reference_cases(case):
    refcases.run_case(case, lambda **kw: oracle.PumMapping(**kw))


def test_clustering_goldens_2d():
    # PartitionOfUnityClusteringTest.cpp:29-62
    pts = [(float(i), float(j)) for i in range(10) for j in range(i, 10)]
    for project, count in ((False, 19), (True, 25)):
        r, _, n = oracle.create_clustering(pts, pts, 2, 0.3, 10, project)
        assert refcases.close(r, 2.2360679774997898)
        assert n == count


def test_clustering_goldens_3d():
    # PartitionOfUnityClusteringTest.cpp:64-94
    pts = [(float(i), float(j), float(k)) for i in range(10) for j in range(i, 10) for k in range(i, 10)]
    for project, count in ((False, 188), (True, 222)):
        r, _, n = oracle.create_clustering(pts, pts, 3, 0.15, 10, project)
        assert refcases.close(r, 1.4142135623730951)
        assert n == count


def test_single_cluster_fallback():
    # PartitionOfUnityMappingTest.cpp:1926-2048: n < 2*vpc -> one cluster, radius 2*longest edge
    r, c, n = oracle.create_clustering(refcases.CUBES_3D, refcases.CUBES_3D, 3, 0.4, 50, False)
    assert n == 1
    assert r == 10.0
    assert np.allclose(c[0], [2.5, 0.5, 0.5])


GLOBAL = [(b, p, poly, c) for b, p, polys in refcases.GLOBAL_BASES for poly in polys for c in refcases.global_cases(b, p, poly)]


@pytest.mark.parametrize("basis,param,poly,case", GLOBAL, ids=[f"{b}-{poly}-{c['name']}" for b, p, poly, c in GLOBAL])
def test_global_reference_cases(basis, param, poly, case):
    refcases.run_global_case(case, lambda **kw: oracle.GlobalMapping(**kw))


DEAD = refcases.dead_axis_cases()


@pytest.mark.parametrize("case", DEAD, ids=[c["name"] for c in DEAD])
def test_global_dead_axis_goldens(case):
    refcases.run_global_case(case, lambda **kw: oracle.GlobalMapping(**kw))


def _explicit_read_pum():
    """tests/serial/just-in-time-mapping/ExplicitReadPUM.{cpp,xml}: MeshTwo = 10 x 10 grid, ten read positions, rbf-pum-direct
    compact-polynomial-c6 support-radius 1, vertices-per-cluster 30, project-to-input off; data of both time steps."""
    grid = np.array([(i * 0.1, j * 0.1) for i in range(10) for j in range(10)], float)
    pos = np.array([(k * 0.2 + 0.05, l * 0.8 + 0.05) for k in range(5) for l in range(2)], float)
    n = np.arange(100, dtype=float)
    data = [(0.5 * n, 1), (-100.0 + 0.1 * n ** 2, 1), (np.stack([0.5 * n, 0.2 * n], 1).reshape(-1), 2),
            (np.stack([0.1 * n * n, 100 - 0.5 * n], 1).reshape(-1), 2)]
    return grid, pos, data


@pytest.mark.parametrize("polynomial", ["separate", "off"])
def test_jit_read_equals_conventional_mapping(polynomial):
    """ExplicitReadPUM.cpp:66-89: mapAndReadData (just-in-time) returns what readData returns after a conventional mapping
    onto a mesh made of the same positions (BOOST_TEST at 1e-9)."""
    grid, pos, data = _explicit_read_pum()
    args = ("consistent", 2, "compact-polynomial-c6", 1.0, polynomial, 30, 0.15, False)
    conv = oracle.PumMapping(*args)
    conv.compute_mapping(grid, pos)
    jit = oracle.PumMapping(*args)
    jit.compute_mapping_jit(grid)
    for vals, dims in data:
        c = jit.initialize_cache(dims)
        jit.update_cache(c, vals)
        assert refcases.close(np.asarray(jit.map_consistent_at(pos, c)).reshape(-1), conv.map(vals, dims))


@pytest.mark.parametrize("polynomial", ["separate", "off"])
def test_jit_write_equals_conventional_mapping(polynomial):
    """tests/serial/just-in-time-mapping/ExplicitWritePUM.{cpp,xml}: writeAndMapData (just-in-time, conservative, Gaussian with
    support radius 1, vertices-per-cluster 30, project-to-input off) onto the 250-vertex grid gives what the conventional
    conservative mapping from a mesh made of the 30 write positions gives (per-element BOOST_TEST), for the scalar and
    3-vector data of both time steps."""
    import math
    grid = np.array([(i * 0.1, j * 0.2, k * 0.2) for i in range(10) for j in range(5) for k in range(5)], float)
    pos = np.array([(k * 0.2 + 0.05, l * 0.8 + 0.05, m * 0.15 + 0.05) for k in range(5) for l in range(2) for m in range(3)], float)
    n = np.arange(30, dtype=float)
    data = [(0.8 * n, 1), (np.stack([12.5 * n, 12.2 * n, -12.2 * n], 1).reshape(-1), 3), (-28.0 + 0.1 * n * n, 1),
            (np.stack([0.28 * n * n, 100 - 0.5 * n, 100 - 0.5 * n ** 3], 1).reshape(-1), 3)]
    args = ("conservative", 3, "gaussian", math.sqrt(-math.log(1e-9)) / 1.0, polynomial, 30, 0.15, False)
    conv = oracle.PumMapping(*args)
    conv.compute_mapping(pos, grid)
    jit = oracle.PumMapping(*args)
    jit.compute_mapping_jit(grid)
    for vals, dims in data:
        expected = np.asarray(conv.map(vals, dims))
        c = jit.initialize_cache(dims)
        jit.map_conservative_at(pos, vals, c)
        out = np.zeros(250 * dims)
        jit.complete_just_in_time_mapping(c, out)
        assert np.max(np.abs(out - expected)) <= 1e-9 * np.max(np.abs(expected))
        if polynomial == "separate":  # the separated polynomial conserves the component sums
            assert refcases.close(out.reshape(-1, dims).sum(0), vals.reshape(-1, dims).sum(0))


INTEGRATION = refcases.gaussian_integration_cases()


@pytest.mark.parametrize("case", INTEGRATION, ids=[c["name"] for c in INTEGRATION])
def test_gaussian_integration_goldens(case):
    """The reference's serial integration tests of the Gaussian rbf-global-direct mapping, at their own tolerances."""
    refcases.run_integration_case(case, lambda **kw: oracle.GlobalMapping(**kw))


PETSC_DISTRIBUTED = refcases.petsc_distributed_cases()


@pytest.mark.parametrize("case", PETSC_DISTRIBUTED, ids=[c["name"] for c in PETSC_DISTRIBUTED])
def test_petsc_distributed_goldens(case):
    """The literals of the PETSc mapping's 4-rank test: the converged iterative solution is the direct one (1e-6)."""
    res = refcases.run_integration_case(case, lambda **kw: oracle.GlobalMapping(**kw))
    assert refcases.close(res.sum(), 36.0)


DISTRIBUTED = refcases.distributed_cases()


@pytest.mark.parametrize("case", DISTRIBUTED, ids=[c["name"] for c in DISTRIBUTED])
def test_global_distributed_goldens(case):
    """The 4-rank gather-scatter test of the reference, as the serial mapping of the union meshes (literals at 1e-9)."""
    refcases.run_global_case(case, lambda **kw: oracle.GlobalMapping(**kw))


# ---------------------------------------------------------------------------------------------- just-in-time mapping
def _jit_boxes():
    def box(x0):
        return [(x0, 0, 0), (x0 + 1, 0, 0), (x0 + 1, 1, 0), (x0, 1, 0), (x0, 0, 1), (x0 + 1, 0, 1), (x0 + 1, 1, 1), (x0, 1, 1)]
    return np.array(box(0) + box(2) + box(4), float)


JIT_VALUES_24 = np.array([1, 2, 2, 1, 3, 6, 6, 3, 3, 4, 4, 3, 9, 12, 12, 9, 5, 6, 6, 5, 15, 18, 18, 15], float)


def test_jit_consistent_3d_with_polynomial():
    """PartitionOfUnityMappingTest.cpp:772-937 (perform3DTestJustInTimeMappingWithPolynomial, instantiated :1908-1909)."""
    m = oracle.PumMapping("consistent", 3, "compact-polynomial-c0", 3.0, "separate", 5, 0.265, False)
    m.compute_mapping_jit(_jit_boxes())
    c = m.initialize_cache(1)
    m.update_cache(c, JIT_VALUES_24)
    at = lambda pts: m.map_consistent_at(np.array(pts, float), c)
    assert at([(0, 0, 0)])[0] == pytest.approx(1.0, rel=1e-12)
    assert at([(3.5, 0.5, 0.5)])[0] == pytest.approx(9.0, rel=1e-12)
    assert at([(2.5, 0.5, 1.0)])[0] == pytest.approx(10.5, rel=1e-12)
    assert at([(0, 0.5, 0.5), (0, 1, 1), (1, 0, 0)]) == pytest.approx([2.0, 3.0, 2.0], rel=1e-12)
    m.update_cache(c, 2 * JIT_VALUES_24)
    assert at([(0, 0.5, 0.5), (0, 1, 1), (1, 0, 0)]) == pytest.approx([4.0, 6.0, 4.0], rel=1e-12)
    assert at([(0.5, 0, 0.5), (0.5, 0.5, 1.0), (0.5, 1.0, 0.0)]) == pytest.approx([6.0, 9.0, 3.0], rel=1e-12)
    # :904-936 the mesh is extended by one more face, the mapping recomputed
    m.clear()
    mesh = np.vstack([_jit_boxes(), [(6, 0, 0), (7, 0, 0), (7, 1, 0), (6, 1, 0)]])
    m.compute_mapping_jit(mesh)
    c = m.initialize_cache(1)
    m.update_cache(c, np.concatenate([JIT_VALUES_24, [7.0, 8.0, 8.0, 7.0]]))
    assert at([(6.5, 1.0, 0.0), (0.5, 0.5, 1.0), (0.5, 1.0, 0.0)]) == pytest.approx([7.5, 4.5, 1.5], rel=1e-12)


def test_jit_consistent_2d_no_polynomial():
    """PartitionOfUnityMappingTest.cpp:939-1098 (perform2DTestJustInTimeMappingNoPolynomial, :1911-1913): goldens at 1e-7."""
    mesh = np.array([(0, 0), (1, 0), (1, 1), (0, 1), (2, 0), (3, 0), (3, 1), (2, 1), (4, 0), (5, 0), (5, 1), (4, 1)], float)
    m = oracle.PumMapping("consistent", 2, "compact-polynomial-c4", 3.0, "off", 5, 0.265, False)
    m.compute_mapping_jit(mesh)
    c = m.initialize_cache(2)
    m.update_cache(c, JIT_VALUES_24)
    at = lambda pts: m.map_consistent_at(np.array(pts, float), c)
    assert at([(0, 0)]) == pytest.approx([1.0, 2.0], rel=1e-12)
    assert at([(5, 1)]) == pytest.approx([15.0, 18.0], rel=1e-12)
    assert at([(0.5, 0.0)]) == pytest.approx([1.71978075, 1.71978075], rel=1e-7)
    assert at([(3.5, 0.5)]) == pytest.approx([10.39684625, 10.48101386], rel=1e-7)
    assert at([(4, 0), (5, 0)]) == pytest.approx([5.0, 6.0, 6.0, 5.0], rel=1e-12)
    m.update_cache(c, 0.5 * JIT_VALUES_24)
    assert at([(4, 0), (5, 0)]) == pytest.approx([2.5, 3.0, 3.0, 2.5], rel=1e-12)


def test_jit_conservative_3d():
    """PartitionOfUnityMappingTest.cpp:1311-1495 (perform3DTestJustInTimeMappingConservative, :1915-1918)."""
    m = oracle.PumMapping("conservative", 3, "compact-polynomial-c2", 3.0, "separate", 5, 0.265, False)
    m.compute_mapping_jit(_jit_boxes())
    c = m.initialize_cache(1)
    out = np.zeros(24)
    m.map_conservative_at([(5.0, 0.0, 1.0)], [4.3], c)
    m.complete_just_in_time_mapping(c, out)
    expected = np.zeros(24)
    expected[21] = 4.3
    assert out.sum() == pytest.approx(4.3, rel=1e-12)
    assert out == pytest.approx(expected, abs=1e-12)
    m.reset_cache(c)
    out[:] = 0
    m.map_conservative_at([(5.0, 0.0, 1.0)], [4.3], c)
    m.map_conservative_at([(3.5, 0.5, 0.5), (4.5, 1.0, 1.0), (0.1, 0.0, 1.0)], [4.3, 17.3, 5.8], c)
    m.complete_just_in_time_mapping(c, out)
    assert out.sum() == pytest.approx(4.3 + 4.3 + 17.3 + 5.8, rel=1e-12)
    assert out[6] == pytest.approx(0.063700286667, rel=1e-7)
    assert out[23] == pytest.approx(9.2744883594, rel=1e-7)
    # :1461-1494 one more vertex, mapping recomputed
    m.clear()
    mesh = np.vstack([_jit_boxes(), [(6, 0, 0)]])
    m.compute_mapping_jit(mesh)
    c = m.initialize_cache(1)
    out = np.zeros(25)
    m.map_conservative_at([(0.5, 0.5, 0.5)], [4.3], c)
    m.map_conservative_at([(6.0, 0.0, 0.0)], [42.0], c)
    m.complete_just_in_time_mapping(c, out)
    assert out.sum() == pytest.approx(46.3, rel=1e-12)
    assert out[24] == pytest.approx(42.0, rel=1e-12)
    assert out[0] == pytest.approx(0.505106848, rel=1e-7)


# ---------------------------------------------------------------------------------------------- committed fixtures (tests/golden)
def _golden():
    import importlib.util, os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, np.load(os.path.join(here, "pum_halton.npz"))


@pytest.mark.parametrize("idx", range(4))
def test_oracle_reproduces_committed_fixtures(idx):
    """tests/golden/pum_halton.npz was written by tests/golden/make_golden.py from this oracle; it must stay reproducible."""
    mod, data = _golden()
    case = mod.CASES[idx]
    support, res, nc, radius = mod.run_case(case)
    assert support == pytest.approx(float(data[case[0] + "/support"]), rel=1e-14)
    assert nc == int(data[case[0] + "/clusters"])
    assert refcases.rel_l2(res, data[case[0] + "/result"]) <= 1e-12


def test_sampled_oracle_matches_full_oracle():
    """oracle.compute_mapping_for_samples (used by tests/golden/make_golden_10m.py to produce the full-size fixture bench.py
    asserts against) is bit-identical to the complete oracle mapping at the sampled vertices."""
    import bench
    import oracle
    n = 12000
    support = bench.support_radius(n)
    inp, out = bench.halton_numpy(n, (2, 3, 5), 1), bench.halton_numpy(n, (7, 11, 13), 1)
    f = np.sin(2 * np.pi * inp[:, 0]) * np.cos(2 * np.pi * inp[:, 1]) + inp[:, 2]
    v3 = np.stack([f, 2 * f + 1, -f], axis=1).reshape(-1)
    samples = np.sort(np.random.default_rng(3).choice(n, 300, replace=False)).astype(np.int32)
    for constraint in ("consistent", "conservative"):
        full = oracle.PumMapping(constraint, 3, bench.BASIS, support, "separate", bench.VPC, bench.OVERLAP, True)
        full.compute_mapping(inp, out)
        part = oracle.PumMapping(constraint, 3, bench.BASIS, support, "separate", bench.VPC, bench.OVERLAP, True)
        part.compute_mapping_for_samples(inp, out, samples)
        assert part.num_clusters == full.num_clusters and part.cluster_radius == full.cluster_radius
        for vals, dims in ((f, 1), (v3, 3)):
            a = full.map(vals, dims).reshape(n, dims)[samples]
            b = part.map(vals, dims).reshape(n, dims)[samples]
            assert np.array_equal(a, b)


def test_full_size_fixture_is_committed():
    import bench
    g = np.load(bench.GOLDEN_10M)
    assert int(g["n"]) == 10_000_000 and g["samples"].shape == (10_000,)
    assert g["consistent_1"].shape == (10_000, 1) and g["conservative_3"].shape == (10_000, 3)
    assert float(g["support_radius"]) == bench.support_radius(10_000_000)
