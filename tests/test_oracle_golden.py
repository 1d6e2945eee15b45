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
