"""Scenarios and golden literals of the reference's own unit tests for the RBF hot path.

Every case is stated as data (meshes, values, expected numbers) so the same table can be run against the
CPU oracle (tests/test_oracle_golden.py, `-m "not gpu"`) and against the CUDA library through its C ABI
(tests/test_gpu_parity.py, `-m gpu`). Sources (relative to /root/reference/src/mapping/tests/):
  PartitionOfUnityMappingTest.cpp:50-770, 1102-1777, 1875-1900   (PUM serial tests, goldens)
  PartitionOfUnityMappingTest.cpp:1779-1873                       (performReferenceTesting meshes)
  PartitionOfUnityClusteringTest.cpp:29-97                        (clustering goldens)
  RadialBasisFctMappingTest.cpp:1416-1546                         (dead-axis goldens)
  RadialBasisFctHelper.hpp:50-695                                 (global mapping linear reproduction)
The reference compares with a relative tolerance of 1e-9 (src/testing/main.cpp:84-87).
"""
from __future__ import annotations

import numpy as np

REL_TOL = 1e-9


def close(a, b, tol=REL_TOL):
    """boost::test_tools::tolerance(1e-9) style: relative to the larger magnitude, with absolute floor."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return bool(np.all(np.abs(a - b) <= tol * np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300) + 1e-13))


# ------------------------------------------------------------------------------------------ meshes
SQUARES_2D = [(0, 0), (1, 0), (1, 1), (0, 1), (2, 0), (3, 0), (3, 1), (2, 1), (4, 0), (5, 0), (5, 1), (4, 1)]
CUBES_3D = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1),
            (2, 0, 0), (3, 0, 0), (3, 1, 0), (2, 1, 0), (2, 0, 1), (3, 0, 1), (3, 1, 1), (2, 1, 1),
            (4, 0, 0), (5, 0, 0), (5, 1, 0), (4, 1, 0), (4, 0, 1), (5, 0, 1), (5, 1, 1), (4, 1, 1)]


def _f(points):
    return np.asarray(points, dtype=np.float64)


# ------------------------------------------------------------------------------------------ PUM cases
# Each step: dict(mapping=kwargs, input=pts, output=pts, values=..., dims=..., checks=[(flat index, expected)],
#                 sum_conserved=bool, extra=callable(out, values) -> None)
# Mapping kwargs follow the reference constructor (PartitionOfUnityMapping.hpp:48-57).

def _pum(constraint, dim, basis, support, vpc, overlap, project=False, polynomial="separate"):
    return dict(constraint=constraint, dim=dim, basis=basis, param=float(support), polynomial=polynomial,
                vertices_per_cluster=vpc, relative_overlap=overlap, project_to_input=project)


def pum_cases():
    cases = []

    # perform2DTestConsistentMapping (:50-317), C0(3), vpc 5, overlap 0.4
    m = _pum("consistent", 2, "compact-polynomial-c0", 3, 5, 0.4)
    vals = [1, 2, 2, 1, 3, 4, 4, 3, 5, 6, 6, 5]
    moving = [((0.0, 0.0), 1.0), ((0.0, 0.5), 1.0), ((0.0, 1.0), 1.0), ((1.0, 0.0), 2.0), ((1.0, 0.5), 2.0),
              ((1.0, 1.0), 2.0), ((0.5, 0.0), 1.5), ((0.5, 0.5), 1.5), ((0.5, 1.0), 1.5)]
    for k, (p, exp) in enumerate(moving):
        out = [p, (3.5, 0.5), (2.5, 0.5), (0, 0), (5, 1)]
        checks = [(0, exp)] + ([(1, 4.5), (2, 3.5)] if k == 0 else [])
        cases.append(dict(name=f"pum2d-consistent-{k}", mapping=m, input=SQUARES_2D, output=out, values=vals, dims=1, checks=checks))

    # perform2DTestConsistentMappingVector (:176-317)
    vals2 = [1, 2, 2, 4, 2, 4, 1, 2, 3, 6, 4, 8, 4, 8, 3, 6, 5, 10, 6, 12, 6, 12, 5, 10]
    for k, (p, exp) in enumerate(moving):
        out = [p, (3.5, 0.5), (2.5, 0.5), (0, 0), (5, 1)]
        checks = [(0, exp), (1, 2 * exp)] + ([(2, 4.5), (3, 9.0), (4, 3.5), (5, 7.0)] if k == 0 else [])
        cases.append(dict(name=f"pum2d-consistent-vector-{k}", mapping=m, input=SQUARES_2D, output=out, values=vals2, dims=2, checks=checks))

    # performTestConsistentMapDeadAxis (:319-378): automatic axis reduction (cond > 1e5)
    for dim, support, overlap in ((2, 6, 0.4), (3, 8, 0.265)):
        inp, val = [], []
        for i in range(20):
            if dim == 2:
                inp += [(0.0, 0.0 + i * dim), (1e-5, 0.0 + i * dim)]
            else:
                inp += [(7.0, 7.0, 40 + i * dim), (7.001, 7.001, 40 + i * dim)]
            val += [i * dim * 1e-5, i * dim * 1e-5 + 2]
        out = [(.1, 1 + 1 + i * dim) if dim == 2 else (7.4, 7.4, 41 + i * dim) for i in range(15)]
        mm = _pum("consistent", dim, "compact-polynomial-c6", support, 5, overlap)
        if dim == 2:
            # reference tolerance here is absolute 0.2 (math::equals(value, ..., 0.2))
            cases.append(dict(name="pum2d-deadaxis", mapping=mm, input=inp, output=out, values=val, dims=1,
                              checks=[(0, 19935.374757794027), (1, 19935.37795225389), (2, 19935.35164938603)], abs_tol=0.2))
        else:
            cases.append(dict(name="pum3d-deadaxis", mapping=mm, input=inp, output=out, values=val, dims=1,
                              checks=[(0, 829.28063055069435), (1, 819.35577218983303), (2, 829.38388713302811)]))

    # perform2DTestConservativeMapping (:380-476)
    mc = _pum("conservative", 2, "compact-polynomial-c0", 3, 5, 0.4)
    inp = [(1.0, 0.0), (3.5, 0.0), (2.5, 0.5), (0, 0), (5, 1)]
    cases.append(dict(name="pum2d-conservative-0", mapping=mc, input=inp, output=SQUARES_2D, values=[10, 0, 0, 10, 0], dims=1,
                      checks=[(0, 10), (1, 10), (2, 0), (3, 0)], sum_conserved=True))
    cases.append(dict(name="pum2d-conservative-1", mapping=mc, input=inp, output=SQUARES_2D, values=[0, 10, 0, 0, 0], dims=1,
                      checks=[], sum_conserved=True, equal_pairs=[(5, 8)]))
    cases.append(dict(name="pum2d-conservative-2", mapping=mc, input=inp, output=SQUARES_2D, values=[5, 10, 7, 3, 4], dims=1,
                      checks=[], sum_conserved=True))
    cases.append(dict(name="pum2d-conservative-3", mapping=mc, input=inp, output=SQUARES_2D, values=[3, 4, 5, 7, 9], dims=1,
                      checks=[(5, 3.2931869752057001)], sum_conserved=True))

    # perform2DTestConservativeMappingVector (:478-623)
    inpv = [(1.0, 0.0), (3.5, 0.5), (2.5, 0.5), (0, 0), (5, 1)]
    cases.append(dict(name="pum2d-conservative-vector-0", mapping=mc, input=inpv, output=SQUARES_2D,
                      values=[1, 2, 2, 4, 2, 4, 1, 2, 3, 6], dims=2, checks=[], component_sums=True))
    cases.append(dict(name="pum2d-conservative-vector-1", mapping=mc, input=inpv, output=SQUARES_2D,
                      values=[1, 0, 12, 0, 2, 0, 1, 0, 3, 0], dims=2, checks=[], component_sums=True))
    cases.append(dict(name="pum2d-conservative-vector-2", mapping=mc, input=inpv, output=SQUARES_2D,
                      values=[0, 2, 0, 4, 0, 8, 0, 7, 0, 0], dims=2, checks=[], component_sums=True))
    cases.append(dict(name="pum2d-conservative-vector-3", mapping=mc, input=inpv, output=SQUARES_2D,
                      values=[10, 0, 0, 0, 0, 0, 0, 0, 0, 27], dims=2, checks=[(2, 10), (21, 27)], component_sums=True))
    cases.append(dict(name="pum2d-conservative-vector-4", mapping=mc, input=inpv, output=SQUARES_2D,
                      values=[3, 6, 2, 4, 12, 24, 1, 2, 3, 6], dims=2,
                      checks=[(12, 3.5933508322619825), (13, 2 * 3.5933508322619825)], sum_conserved=True))

    # perform3DTestConsistentMapping (:625-768), C0(3), vpc 5, overlap 0.265
    m3 = _pum("consistent", 3, "compact-polynomial-c0", 3, 5, 0.265)
    vals3 = [1, 2, 2, 1, 3, 6, 6, 3, 3, 4, 4, 3, 9, 12, 12, 9, 5, 6, 6, 5, 15, 18, 18, 15]
    moving3 = [((0, 0, 0), 1.0), ((0, .5, .5), 2.0), ((0, 1, 1), 3.0), ((1, 0, 0), 2.0), ((1, .5, .5), 4.0),
               ((1, 1, 1), 6.0), ((.5, 0, .5), 3.0), ((.5, .5, 1), 4.5), ((.5, 1, 0), 1.5)]
    for k, (p, exp) in enumerate(moving3):
        out = [p, (3.5, 0.5, 0.5), (2.5, 0.5, 1.0), (0, 0, 0.5), (5, 1, 0.5)]
        checks = [(0, exp)] + ([(1, 9.0), (2, 10.5)] if k == 0 else [])
        cases.append(dict(name=f"pum3d-consistent-{k}", mapping=m3, input=CUBES_3D, output=out, values=vals3, dims=1, checks=checks))

    # perform3DTestConsistentMappingVector (:1102-1197)
    out3 = [(0, 0, 0), (3.5, 0.5, 0.5), (2.5, 0.5, 1.0), (0, 0, 0.5), (5, 1, 0.5)]
    cases.append(dict(name="pum3d-consistent-vector-const", mapping=m3, input=CUBES_3D, output=out3,
                      values=[7, 38, 19] * 24, dims=3, checks=[(0, 7), (1, 38), (2, 19), (9, 7), (10, 38), (11, 19)]))
    vq = []
    for i in range(24):
        vq += [(i * 3) ** 2, 2 * (i * 3) ** 2, 3 * (i * 3) ** 2]
    cases.append(dict(name="pum3d-consistent-vector-quadratic", mapping=m3, input=CUBES_3D, output=out3, values=vq, dims=3,
                      checks=[(12, 3673.7308684337013), (13, 2 * 3673.7308684337013), (14, 3 * 3673.7308684337013)]))

    # perform3DTestConservativeMapping (:1199-1309)
    mc3 = _pum("conservative", 3, "compact-polynomial-c0", 3, 5, 0.265)
    inp3 = [(0, 0, 0), (3.5, 0.5, 0.5), (2.5, 0.5, 1.0), (1, 1, 1), (5, 1, 0.5)]
    cases.append(dict(name="pum3d-conservative-0", mapping=mc3, input=inp3, output=CUBES_3D, values=[1, 2, 2, 1, 3], dims=1,
                      checks=[], sum_conserved=True))
    cases.append(dict(name="pum3d-conservative-1", mapping=mc3, input=inp3, output=CUBES_3D, values=[12, 5, 7, 8, 9], dims=1,
                      checks=[(6, 8), (0, 12)], sum_conserved=True))
    cases.append(dict(name="pum3d-conservative-2", mapping=mc3, input=inp3, output=CUBES_3D, values=[0, 50, 107, 108, 48], dims=1,
                      checks=[(0, 0.0)], sum_conserved=True, greater_pairs=[(6, 1), (13, 9)]))
    cases.append(dict(name="pum3d-conservative-3", mapping=mc3, input=inp3, output=CUBES_3D, values=[0, 100, 0, 0, 0], dims=1,
                      checks=[(9, 15.470584170385226)], sum_conserved=True,
                      equal_pairs=[(9, 13), (9, 10), (16, 19), (20, 23), (16, 20)]))

    # perform3DTestConservativeMappingVector (:1640-1777)
    for c, (v, s) in enumerate(((7, 35), (27, 135), (3, 15))):
        vv = []
        for i in range(5):
            e = [0, 0, 0]
            e[c] = v
            vv += e
        cases.append(dict(name=f"pum3d-conservative-vector-{c}", mapping=mc3, input=inp3, output=CUBES_3D, values=vv, dims=3,
                          checks=[], component_sums=True, total=s))
    vc = []
    for i in range(5):
        vc += [(i * 3) ** 3, 5 * (i * 3) ** 3, 10 * (i * 3) ** 3]
    cases.append(dict(name="pum3d-conservative-vector-cubic", mapping=mc3, input=inp3, output=CUBES_3D, values=vc, dims=3,
                      checks=[(55, 4087.8100404079933)], sum_conserved=True, component_ratio=(1, 5, 10)))
    return cases


def run_case(case, make_mapping):
    """make_mapping(**kwargs) -> object with compute_mapping(in_xyz, out_xyz), map(values, dims) -> flat array."""
    m = make_mapping(**case["mapping"])
    inp, out = _f(case["input"]), _f(case["output"])
    vals = np.asarray(case["values"], dtype=np.float64)
    dims = case["dims"]
    m.compute_mapping(inp, out)
    res = np.asarray(m.map(vals, dims))
    assert res.shape == (out.shape[0] * dims,)
    for idx, exp in case["checks"]:
        if "abs_tol" in case:
            assert abs(res[idx] - exp) <= case["abs_tol"], (case["name"], idx, res[idx], exp)
        else:
            assert close(res[idx], exp), (case["name"], idx, repr(res[idx]), exp)
    if case.get("sum_conserved"):
        assert close(res.sum(), vals.sum()), (case["name"], res.sum(), vals.sum())
    if case.get("component_sums"):
        for c in range(dims):
            assert close(res[c::dims].sum(), vals[c::dims].sum()), (case["name"], c, res[c::dims].sum(), vals[c::dims].sum())
            if vals[c::dims].sum() == 0:
                assert np.all(np.abs(res[c::dims]) <= 1e-13), (case["name"], c)
    if "total" in case:
        assert close(res.sum(), case["total"])
    for a, b in case.get("equal_pairs", []):
        assert close(res[a], res[b]), (case["name"], a, b, res[a], res[b])
    for a, b in case.get("greater_pairs", []):
        assert res[a] > res[b], (case["name"], a, b)
    if "component_ratio" in case:
        r = case["component_ratio"]
        for i in range(out.shape[0]):
            if res[i * dims] > 1e-10:
                for c in range(1, dims):
                    assert close(res[i * dims + c], r[c] * res[i * dims]), (case["name"], i, c)
    return res


# ------------------------------------------------------------------------------------------ reference-testing meshes
def reference_testing_meshes(dim):
    """performReferenceTesting (PartitionOfUnityMappingTest.cpp:1779-1873): grid -> offset grid, ramp data."""
    if dim == 2:
        n = 10
        inp = [(i / (n - 1), j / (n - 1)) for i in range(n) for j in range(n)]
        n = 9
        out = [((i + 0.2) / n, (j + 0.5) / n) for i in range(n) for j in range(n)]
    else:
        n = 7
        inp = [(i / (n - 1), j / (n - 1), k / (n - 1)) for i in range(n) for j in range(n) for k in range(n)]
        n = 8
        out = [((i + 0.2) / n, (j + 0.3) / n, (k + 0.35) / n) for i in range(n) for j in range(n) for k in range(n)]
    return _f(inp), _f(out)


def reference_testing_values(n_vertices, n_components):
    base = np.arange(n_vertices, dtype=np.float64) / float(n_vertices - 1)
    vals = np.empty(n_vertices * n_components)
    for c in range(n_components):
        vals[c::n_components] = (2 * c + 1) * base
    return vals


# ------------------------------------------------------------------------------------------ synthetic meshes
def halton(n, bases, skip=1):
    """Deterministic Halton points in the unit cube (benchmarks/helper.hpp:6-41 uses bases 2/3/5)."""
    out = np.empty((n, len(bases)))
    idx = np.arange(skip, n + skip, dtype=np.int64)
    for d, b in enumerate(bases):
        f = 1.0
        r = np.zeros(n)
        i = idx.copy()
        while np.any(i > 0):
            f /= b
            r += f * (i % b)
            i //= b
        out[:, d] = r
    return out


def field(xyz):
    """f(x) = sin(2 pi x) cos(2 pi y) + z   """
    x = np.asarray(xyz)
    z = x[:, 2] if x.shape[1] > 2 else 0.0
    return np.sin(2 * np.pi * x[:, 0]) * np.cos(2 * np.pi * x[:, 1]) + z


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# ------------------------------------------------------------------------------------------ global-direct cases
# (basis name, constructor parameter, polynomial modes tested) as in RadialBasisFctMappingTest.cpp:1318-1412
GLOBAL_BASES = [
    ("thin-plate-splines", 0.0, ("on", "separate")),
    ("multiquadrics", 1e-3, ("on", "separate")),
    ("inverse-multiquadrics", 1e-3, ("separate",)),
    ("volume-splines", 0.0, ("on", "separate")),
    ("gaussian", 1.0, ("separate",)),
    ("compact-tps-c2", 1.2, ("separate",)),
    ("compact-polynomial-c0", 1.2, ("separate",)),
    ("compact-polynomial-c2", 1.2, ("separate",)),
    ("compact-polynomial-c4", 1.2, ("separate",)),
    ("compact-polynomial-c6", 1.2, ("separate",)),
    ("compact-polynomial-c8", 1.2, ("separate",)),
]


def global_cases(basis, param, polynomial):
    """RadialBasisFctHelper.hpp:50-695 — linear reproduction / conservation scenarios of the global mapping."""
    def mk(constraint, dim):
        return dict(constraint=constraint, dim=dim, basis=basis, param=param, dead_axis=(False, False, False), polynomial=polynomial)
    cases = []
    sq = [(0, 0), (1, 0), (1, 1), (0, 1)]
    # perform2DTestConsistentMapping :50-145
    pts = [((0, 0), 1.0), ((0, .5), 1.0), ((0, 1), 1.0), ((1, 0), 2.0), ((1, .5), 2.0), ((1, 1), 2.0), ((.5, 0), 1.5), ((.5, .5), 1.5), ((.5, 1), 1.5)]
    for k, (p, e) in enumerate(pts):
        cases.append(dict(name=f"g2d-consistent-{k}", mapping=mk("consistent", 2), input=sq, output=[p], values=[1, 2, 2, 1], dims=1, checks=[(0, e)]))
    # perform2DTestConsistentMappingVector :147-260
    for k, (p, e) in enumerate(pts):
        cases.append(dict(name=f"g2d-consistent-vector-{k}", mapping=mk("consistent", 2), input=sq, output=[p],
                          values=[1, 4, 2, 5, 2, 5, 1, 4], dims=2, checks=[(0, e), (1, e + 3.0)]))
    # perform3DTestConsistentMapping :262-395
    cube = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
    pts3 = [((0, 0, 0), 1.0), ((0, .5, 0), 1.0), ((.5, .5, 0), 1.0), ((1, 1, 0), 1.0), ((0, 0, 1), 2.0), ((1, 1, 1), 2.0),
            ((.5, .5, 1), 2.0), ((0, 0, .5), 1.5), ((1, 0, .5), 1.5), ((.5, .5, .5), 1.5)]
    for k, (p, e) in enumerate(pts3):
        cases.append(dict(name=f"g3d-consistent-{k}", mapping=mk("consistent", 3), input=cube, output=[p],
                          values=[1, 1, 1, 1, 2, 2, 2, 2], dims=1, checks=[(0, e)]))
    # perform2DTestConservativeMapping :505-573 (tolerance 1e-6)
    cons = [(((.5, 0), (.5, 1)), [.5, .5, 1, 1]), (((0, .5), (1, .5)), [.5, 1, 1, .5]),
            (((0, 1), (1, 0)), [0, 2, 0, 1]), (((0, 0), (1, 1)), [1, 0, 2, 0])]
    for k, (inp, e) in enumerate(cons):
        cases.append(dict(name=f"g2d-conservative-{k}", mapping=mk("conservative", 2), input=list(inp), output=sq, values=[1, 2], dims=1,
                          checks=list(enumerate(e)), tol=1e-6))
    cases.append(dict(name="g2d-conservative-sum", mapping=mk("conservative", 2), input=[(.4, .5), (.6, .5)], output=sq, values=[1, 2], dims=1,
                      checks=[], sum_conserved=True))
    # perform3DTestConservativeMapping :650-694
    cube2 = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    cases.append(dict(name="g3d-conservative-sum", mapping=mk("conservative", 3), input=[(.5, 0, 0), (.5, 1, 0)], output=cube2, values=[1, 2], dims=1,
                      checks=[], sum_conserved=True))
    return cases


def dead_axis_cases():
    """RadialBasisFctMappingTest.cpp:1416-1546 — thin-plate splines with y dead; golden literals."""
    cases = []
    in2 = [(0, 1), (1, 1), (2, 1), (3, 1)]
    out2 = [(0, 1), (3, 1), (1.3, 1), (5, 1)]
    g2 = {("consistent", "off"): (2, 2.0522549299731567), ("consistent", "separate"): (2, 2.0896514371485777),
          ("consistent", "on"): (2, 2.1180354377884774), ("conservative", "off"): (1, 1.8471144693068295),
          ("conservative", "separate"): (1, 1.8236736422730249), ("conservative", "on"): (1, 1.7587181970483183)}
    for (constraint, poly), (idx, val) in g2.items():
        cases.append(dict(name=f"deadaxis2d-{constraint}-{poly}",
                          mapping=dict(constraint=constraint, dim=2, basis="thin-plate-splines", param=0.0, dead_axis=(False, True, False), polynomial=poly),
                          input=in2, output=out2, values=[1, 2, 2, 1], dims=1, checks=[(idx, val)], tol=1e-7))
    in3 = [(0, 3, 0), (1, 3, 0), (0, 3, 1), (1, 3, 1)]
    out3 = [(0, 2.9, 0), (.8, 2.9, .1), (.1, 2.9, .9), (1.1, 2.9, 1.1)]
    g3 = {("consistent", "off"): ([1.0, -0.454450524334, 0.99146426249, 6.98958304876], 1e-7),
          ("consistent", "on"): ([1.0, 2.0, 2.9, 4.3], 1e-9), ("consistent", "separate"): ([1.0, 2.0, 2.9, 4.3], 1e-9),
          ("conservative", "off"): ([1.17251596926, 4.10368825944, 3.56931954192, 3.40160932341], 1e-6),
          ("conservative", "on"): ([0.856701171969, 2.38947124326, 3.34078733786, 3.41304024691], 1e-6),
          ("conservative", "separate"): ([0.380480856704, 2.83529451713, 3.73088270249, 3.05334192368], 1e-6)}
    for (constraint, poly), (vals, tol) in g3.items():
        cases.append(dict(name=f"deadaxis3d-{constraint}-{poly}",
                          mapping=dict(constraint=constraint, dim=3, basis="thin-plate-splines", param=0.0, dead_axis=(False, True, False), polynomial=poly),
                          input=in3, output=out3, values=[1, 2, 3, 4], dims=1, checks=list(enumerate(vals)), tol=tol))
    return cases


def distributed_cases():
    """RadialBasisFctMappingTest.cpp:969-1034 (DistributedConservative2DV4): the reference gathers the distributed meshes
    on rank 0, solves the global system there and scatters (RadialBasisFctMapping.hpp:116-149, 199-308), so the expected
    values of the 4-rank test are those of the serial mapping between the unions of the rank-local meshes - full-precision
    literals for a conservative thin-plate-spline mapping onto a smaller mesh."""
    inp = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1)]
    out = [(0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1)]
    golden = [3.1307987056295525, 4.0734031184906971, 3.0620533966233006, 4.4582564956060873, 5.8784343572772633,
              7.4683403859032156, 7.9287135404698859]  # :1009-1029, sum ~36
    return [dict(name="distributed-conservative-2d-v4",
                 mapping=dict(constraint="conservative", dim=2, basis="thin-plate-splines", param=0.0, dead_axis=(False, False, False), polynomial="separate"),
                 input=inp, output=out, values=[1, 2, 3, 4, 5, 6, 7, 8], dims=1, checks=list(enumerate(golden)), sum_conserved=True)]


def gaussian_integration_cases():
    """tests/serial/mapping-rbf-gaussian (helpers.cpp:12-190 + the four XML files): rbf-global-direct, Gaussian given by a
    support radius or the equivalent shape parameter (MappingConfiguration.cpp:512-521: shape = sqrt(-ln 1e-9) / support),
    polynomial separate, z-dead, MeshOne (12 vertices at z = 0.3) -> MeshTwo (3 vertices). Each check carries the relative
    tolerance of the reference's BOOST_TEST (the scalar cases are badly conditioned: "Eigen 3.3.7 ... slightly different")."""
    import math
    z = 0.3
    one = [(0, 0, z), (1, 0, z), (1, 1, z), (0, 1, z), (2, 0, z), (3, 0, z), (3, 1, z), (2, 1, z), (4, 0, z), (5, 0, z), (5, 1, z), (4, 1, z)]
    two = [(0, 0, z), (0.5, 0.5, z), (3.5, 0.5, z)]
    cut = math.sqrt(-math.log(1e-9))

    def mapping(constraint, shape):
        return dict(constraint=constraint, dim=3, basis="gaussian", param=shape, dead_axis=(False, False, True), polynomial="separate")

    def linear(mesh):  # helpers.cpp:55-68
        return [0.1 * sum(v) + c for v in mesh for c in (0, 1, 2)]

    cases = []
    squares = [float((i + 1) ** 2) for i in range(12)]  # helpers.cpp:125-131
    for name, shape in (("shape-parameter", 2.2761406940777196), ("support-radius", cut / 2.0)):
        cases.append(dict(name=f"gaussian-integration-{name}", mapping=mapping("consistent", shape), input=one, output=two, values=squares, dims=1,
                          checks=[(0, 1.0000000014191541, 1e-8), (1, 7.30892688709867, 3e-2), (2, 77.5938805368033, 2e-3)]))
    cases.append(dict(name="gaussian-integration-vectorial-consistent", mapping=mapping("consistent", cut / 100.0), input=one, output=two,
                      values=linear(one), dims=3, checks=[(i, v, 1e-5) for i, v in enumerate(linear(two))]))
    cons = [-0.6, -0.6, -0.6, 0.853333, 4.85333, 8.85333, 3.70667, 11.7067, 19.7067]  # helpers.cpp:103
    cases.append(dict(name="gaussian-integration-vectorial-conservative", mapping=mapping("conservative", cut / 100.0), input=one, output=two,
                      values=linear(one), dims=3, checks=[(i, v, 1e-5) for i, v in enumerate(cons)]))
    return cases


def petsc_distributed_cases():
    """PetRadialBasisFctMappingTest.cpp:850-909 (DistributedConservative2DV4 of the PETSc mapping, 4 ranks, tolerance 1e-6):
    Gaussian(4.0), conservative, separated polynomial (the PETSc mapping's default), the meshes of distributed_cases()."""
    inp = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1)]
    out = [(0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1)]
    golden = [2.4285714526861519, 3.61905, 4.14286, 5.333333295, 5.85714, 7.047619, 7.571428]  # :888-905, sum ~36
    return [dict(name="petsc-distributed-conservative-2d-v4",
                 mapping=dict(constraint="conservative", dim=2, basis="gaussian", param=4.0, dead_axis=(False, False, False), polynomial="separate"),
                 input=inp, output=out, values=[1, 2, 3, 4, 5, 6, 7, 8], dims=1, checks=[(i, v, 1e-6) for i, v in enumerate(golden)])]


def run_integration_case(case, make_mapping):
    """Checks with a relative tolerance per value, boost::test_tools::tolerance style."""
    m = make_mapping(**case["mapping"])
    m.compute_mapping(_f(case["input"]), _f(case["output"]))
    res = np.asarray(m.map(np.asarray(case["values"], dtype=np.float64), case["dims"]))
    for idx, exp, tol in case["checks"]:
        assert abs(res[idx] - exp) <= tol * max(abs(res[idx]), abs(exp)), (case["name"], idx, repr(res[idx]), exp, tol)
    return res


def run_global_case(case, make_mapping):
    m = make_mapping(**case["mapping"])
    inp, out = _f(case["input"]), _f(case["output"])
    vals = np.asarray(case["values"], dtype=np.float64)
    m.compute_mapping(inp, out)
    res = np.asarray(m.map(vals, case["dims"]))
    tol = case.get("tol", REL_TOL)
    for idx, exp in case["checks"]:
        # testing::equals(a, b, tol): |a-b| <= tol (absolute) for the 1e-6/1e-7 literals; 1e-9 relative otherwise
        if tol > REL_TOL:
            assert abs(res[idx] - exp) <= tol, (case["name"], idx, repr(res[idx]), exp)
        else:
            assert close(res[idx], exp, tol), (case["name"], idx, repr(res[idx]), exp)
    if case.get("sum_conserved"):
        assert close(res.sum(), vals.sum()), (case["name"], res.sum(), vals.sum())
    return res
