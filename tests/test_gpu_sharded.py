"""`-m gpu`: ONE rbf-pum-direct mapping sharded over a group of ranks (include/rbfmap.h) against the
single-device mapping and the CPU oracle. Tolerance 1e-10 relative L2 (north_star).

* duplicate mode needs no collective, so W "virtual ranks" (W handles of this process, all on device 0, joined without an
  NCCL id) exercise the complete sharded code path on a single-GPU box: slab cuts, cluster selection, slab-filtered
  membership and weights, per-slab unassigned check;
* exchange mode (and the conservative constraint) all-reduce over NCCL and need one GPU per rank: those tests launch
  scripts/sharded_pum.py under torchrun and are skipped when fewer than two GPUs are visible (`gpurun --gpus 2`).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import refcases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL_L2_TOL = 1e-10


@pytest.fixture(scope="module")
def pb():
    import precice_b200
    if precice_b200.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need a B200 (no CPU fallback exists)")
    return precice_b200


def _meshes(n, dim=3):
    if dim == 3:
        return refcases.halton(n, (2, 3, 5)), refcases.halton(n, (7, 11, 13))
    a, b = refcases.halton(n, (2, 3, 5)), refcases.halton(n, (7, 11, 13))
    a[:, 2] = 0.0
    b[:, 2] = 0.0
    return a, b


def _field(x, dims):
    f = np.sin(2 * np.pi * x[:, 0]) * np.cos(2 * np.pi * x[:, 1]) + x[:, 2]
    return f if dims == 1 else np.stack([f, 2 * f + 1, -f], axis=1).reshape(-1)


def _support(n, vpc, dim):
    if dim == 3:
        return 2.0 * (3.0 * vpc / (4.0 * np.pi * n)) ** (1.0 / 3.0)
    return 2.0 * (vpc / (np.pi * n)) ** 0.5


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("dim,dims,project", [(3, 1, True), (3, 3, False), (2, 1, True)])
def test_pum_sharded_duplicate_virtual_ranks(pb, world, dim, dims, project):
    import oracle
    n, vpc = 6000, 30
    inp, out = _meshes(n, dim)
    vals = _field(inp, dims)
    args = ("consistent", dim, "compact-polynomial-c6", _support(n, vpc, dim), "separate", vpc, 0.15, project)
    single = pb.PartitionOfUnityMapping(*args)
    single.computeMapping(inp, out)
    ref = single.map(vals, dims)
    st0 = single.stats()
    assert st0["shard_axis"] == -1 and st0["n_clusters_global"] == st0["n_clusters"]
    _, centers0, *_ = single.clusters()

    total = np.zeros_like(ref)
    owners = np.zeros(n, dtype=int)
    processed, seen = 0, set()
    for r in range(world):
        m = pb.PartitionOfUnityMapping(*args, shard_mode="duplicate")
        m.distributedInit(None, r, world)
        m.computeMapping(inp, out)
        part = m.map(vals, dims)
        st = m.stats()
        assert st["n_clusters_global"] == -1  # unknown without communicator: every rank only clusters its own region
        ax, lo, hi = st["shard_axis"], st["shard_lo"], st["shard_hi"]
        assert 0 <= ax < dim
        own = (out[:, ax] >= lo) & (out[:, ax] < hi)
        owners += own
        # the rank fills exactly the output vertices it owns, and those already carry the final value
        assert np.all(part.reshape(n, dims)[~own] == 0.0)
        assert refcases.rel_l2(part.reshape(n, dims)[own], ref.reshape(n, dims)[own]) <= 1e-13
        total += part
        processed += st["n_clusters"]
        if st["n_clusters"]:
            _, c, in_off, out_off, in_ids, out_ids = m.clusters()
            seen.update(map(tuple, c))
            assert set(out_ids.tolist()) <= set(np.flatnonzero(own).tolist())
    assert np.all(owners == 1)  # the slabs partition the output mesh
    assert seen == set(map(tuple, centers0))  # every cluster is processed somewhere, none is invented
    assert st0["n_clusters"] <= processed <= 4 * st0["n_clusters"]
    assert refcases.rel_l2(total, ref) <= 1e-13
    orc = oracle.PumMapping(*args)
    orc.compute_mapping(inp, out)
    assert refcases.rel_l2(total, orc.map(vals, dims)) <= REL_L2_TOL


def test_pum_sharded_more_ranks_than_slabs(pb):
    """Single-cluster fallback (CreateClustering.hpp:352-359) and tiny meshes: the surplus ranks own nothing and still work."""
    inp, out = _meshes(40)
    vals = _field(inp, 1)
    args = ("consistent", 3, "compact-polynomial-c6", 3.0, "separate", 50, 0.15, True)
    single = pb.PartitionOfUnityMapping(*args)
    single.computeMapping(inp, out)
    ref = single.map(vals, 1)
    total = np.zeros_like(ref)
    for r in range(4):
        m = pb.PartitionOfUnityMapping(*args)
        m.distributedInit(None, r, 4)
        m.computeMapping(inp, out)
        total += m.map(vals, 1)
        assert m.stats()["n_clusters"] == (1 if r == 0 else 0)
    assert refcases.rel_l2(total, ref) <= 1e-14


def test_pum_sharded_unassigned_vertex_is_reported_by_its_owner(pb):
    """PartitionOfUnityMapping.hpp:314-318 on a sharded mapping: the rank that owns the stray vertex raises the error."""
    n, vpc = 4000, 30
    inp, out = _meshes(n)
    out = np.concatenate([out, [[3.0, 0.5, 0.5]]])
    args = ("consistent", 3, "compact-polynomial-c6", _support(n, vpc, 3), "separate", vpc, 0.15, True)
    raised = 0
    for r in range(2):
        m = pb.PartitionOfUnityMapping(*args)
        m.distributedInit(None, r, 2)
        try:
            m.computeMapping(inp, out)
        except pb.MappingError as e:
            assert e.status == "UNASSIGNED" and "could not be assigned to any cluster" in str(e)
            raised += 1
    assert raised == 1


def test_group_without_communicator_is_limited_to_duplicate_consistent(pb):
    m = pb.PartitionOfUnityMapping("consistent", 3, "compact-polynomial-c6", 0.3, shard_mode="exchange")
    with pytest.raises(pb.MappingError) as e:
        m.distributedInit(None, 0, 2)
    assert e.value.status == "INVALID"
    m = pb.PartitionOfUnityMapping("conservative", 3, "compact-polynomial-c6", 0.3)
    with pytest.raises(pb.MappingError):
        m.distributedInit(None, 0, 2)
    m = pb.RadialBasisFctMapping("consistent", 3, "gaussian", 5.0, solver="iterative")
    with pytest.raises(pb.MappingError):
        m.distributedInit(None, 0, 2)


def test_map_sliced_on_one_device_and_without_communicator(pb):
    """rbfmap_map_sliced: on a single device the slice is the whole result; a group without communicator cannot complete
    the slices of a duplicate-mode mapping."""
    n, vpc = 3000, 30
    inp, out = _meshes(n)
    vals = _field(inp, 3)
    args = ("consistent", 3, "compact-polynomial-c6", _support(n, vpc, 3), "separate", vpc, 0.15, True)
    m = pb.PartitionOfUnityMapping(*args)
    m.computeMapping(inp, out)
    res, b, e = m.mapSliced(vals, 3)
    assert (b, e) == (0, out.shape[0]) and refcases.rel_l2(res, m.map(vals, 3)) <= 1e-14  # the blend uses atomics
    v = pb.PartitionOfUnityMapping(*args)
    v.distributedInit(None, 1, 2)
    v.computeMapping(inp, out)
    with pytest.raises(pb.MappingError) as err:
        v.mapSliced(vals, 3)
    assert err.value.status == "UNSUPPORTED"


def _torchrun(nproc, *script_args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "sharded_pum.py"), *script_args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


def test_pum_sharded_two_gpus(pb):
    """Both shard modes, consistent + conservative, scalar + 3-vector, host and device pointers, over NCCL on two GPUs:
    sharded == single device == oracle at 1e-10 (scripts/sharded_pum.py asserts, this test checks its verdict line)."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with `gpurun --gpus 2`)")
    out = _torchrun(2, "--vertices", "20000", "--oracle")
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARDED PUM OK" in out.stdout, out.stdout[-3000:]


# ---------------------------------------------------------------------------------------------- re-partitioning hooks
def _box_tags(remote, local, margin, dim):
    lo, hi = local.min(0) - margin, local.max(0) + margin
    inside = np.ones(len(remote), bool)
    for d in range(dim):
        inside &= ~((remote[:, d] < lo[d]) | (remote[:, d] > hi[d]))
    return inside.astype(np.uint8)


@pytest.mark.parametrize("constraint", ["consistent", "conservative"])
def test_pum_tag_mesh_first_round(pb, constraint):
    """PartitionOfUnityMapping.hpp:479-533: bounding box of the local mesh expanded by 2 x the cluster radius, the radius
    estimated on the unfiltered remote mesh with the local box (impl/CreateClustering.hpp:223-278)."""
    import oracle
    remote = refcases.halton(30000, (2, 3, 5))
    local = 0.3 + 0.25 * refcases.halton(2000, (7, 11, 13))  # this rank's partition: a sub-box of the remote mesh
    m = pb.PartitionOfUnityMapping(constraint, 3, "compact-polynomial-c6", 0.3, "separate", 40, 0.15, True)
    inp, out = (remote, local) if constraint == "consistent" else (local, remote)
    radius = m.estimateClusterRadius(remote, local.min(0), local.max(0))
    assert radius == oracle.estimate_cluster_radius(remote, 3, 40, local.min(0), local.max(0))
    tags = m.tagMeshFirstRound(inp, out)
    assert tags.shape == (30000,) and np.array_equal(tags, _box_tags(remote, local, 2 * radius, 3))
    assert 0 < tags.sum() < 30000
    # tags are only ever set (Vertex::tag), and the second round leaves them alone (:535-539)
    pre = np.zeros(30000, np.uint8)
    pre[:10] = 1
    assert np.array_equal(m.tagMeshFirstRound(inp, out, pre), np.maximum(tags, pre))
    before = tags.copy()
    m.tagMeshSecondRound(remote, np.ones(30000, np.uint8), tags)
    assert np.array_equal(tags, before)
    # an empty local mesh tags nothing (:491-492)
    empty = np.zeros((0, 3))
    none = m.tagMeshFirstRound(*((remote, empty) if constraint == "consistent" else (empty, remote)))
    assert none.sum() == 0


def test_global_tag_mesh_rounds(pb):
    """RadialBasisFctBaseMapping.hpp:124-190."""
    remote = refcases.halton(20000, (2, 3, 5))
    local = 0.4 + 0.2 * refcases.halton(1500, (7, 11, 13))
    m = pb.RadialBasisFctMapping("consistent", 3, "compact-polynomial-c2", 0.07)
    tags = m.tagMeshFirstRound(remote, local)
    assert np.array_equal(tags, _box_tags(remote, local, 0.07, 3))
    owner = (tags == 1) & (remote[:, 0] < 0.5)
    second = m.tagMeshSecondRound(remote, owner.astype(np.uint8), tags.copy())
    assert np.array_equal(second, np.maximum(tags, _box_tags(remote, remote[owner], 0.07, 3)))
    # conservative: the remote mesh is output(); 2-D ignores z
    r2, l2 = remote.copy(), local.copy()
    r2[:, 2] = 0.0
    l2[:, 2] = 0.0
    mc = pb.RadialBasisFctMapping("conservative", 2, "compact-polynomial-c2", 0.07)
    assert np.array_equal(mc.tagMeshFirstRound(l2, r2), _box_tags(r2, l2, 0.07, 2))
    # global support: everything is tagged (:153-155), the second round changes nothing (:167-168)
    mt = pb.RadialBasisFctMapping("consistent", 3, "thin-plate-splines")
    assert mt.tagMeshFirstRound(remote, local).all()
    z = np.zeros(20000, np.uint8)
    assert mt.tagMeshSecondRound(remote, np.ones(20000, np.uint8), z).sum() == 0


@pytest.mark.parametrize("long_axis", [1, 2])
def test_pum_sharded_slab_axis_follows_the_lattice(pb, long_axis):
    """The slabs are cut along the lattice axis with the most layers: meshes elongated along y / z are cut along y / z (the
    lattice index is x-major, so only axis 0 gives contiguous index ranges - the other axes exercise the general path)."""
    n, vpc = 8000, 30
    inp, out = _meshes(n)
    inp[:, long_axis] *= 3.0
    out[:, long_axis] *= 3.0
    vals = _field(inp, 1)
    support = 2.0 * (3.0 * 3.0 * vpc / (4.0 * np.pi * n)) ** (1.0 / 3.0)
    args = ("consistent", 3, "compact-polynomial-c6", support, "separate", vpc, 0.15, True)
    single = pb.PartitionOfUnityMapping(*args)
    single.computeMapping(inp, out)
    ref = single.map(vals, 1)
    _, centers0, *_ = single.clusters()
    total, seen, owners = np.zeros_like(ref), set(), np.zeros(n, dtype=int)
    for r in range(4):
        m = pb.PartitionOfUnityMapping(*args)
        m.distributedInit(None, r, 4)
        m.computeMapping(inp, out)
        st = m.stats()
        assert st["shard_axis"] == long_axis
        owners += (out[:, long_axis] >= st["shard_lo"]) & (out[:, long_axis] < st["shard_hi"])
        total += m.map(vals, 1)
        if st["n_clusters"]:
            seen.update(map(tuple, m.clusters()[1]))
    assert np.all(owners == 1) and seen == set(map(tuple, centers0))
    assert refcases.rel_l2(total, ref) <= 1e-13
