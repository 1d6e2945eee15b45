"""`-m "not gpu"`: the N > 1 path on the CPU, world_size 2 on the gloo backend.

No kernel can run here, so what is exercised is everything around them: the decomposition the sharded rbf-pum-direct
uses (include/rbfmap.h) — slab cuts from the PRODUCT's host helper `rbfmap_partition_slabs` in librbfmap.so,
cluster selection by slab (duplicate mode: every cluster whose sphere reaches into the slab, output vertices owned by
slab; exchange mode: clusters owned by centre, partial blends and weight sums all-reduced) — carried out on the oracle's
clusters with numpy per-cluster algebra, reduced over gloo and compared with the serial oracle at 1e-10; plus bench.py's
reference arm. The device implementation of the same decomposition is tested by tests/test_gpu_sharded.py.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_BINS = 4096


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _c6(r, support):
    p = np.minimum(r / support, 1.0)
    return (1 - p) ** 8 * (32 * p ** 3 + 25 * p ** 2 + 8 * p + 1)


def _c2(p):
    p = np.minimum(p, 1.0)
    return (1 - p) ** 4 * (4 * p + 1)


def _cluster_interpolant(inp, out_pts, f, support):
    """RadialBasisFctSolver::solveConsistent with a separated polynomial (RadialBasisFctSolver.hpp:441-468), numpy."""
    d = np.linalg.norm(inp[:, None, :] - inp[None, :, :], axis=2)
    Cm = _c6(d, support)
    Q = np.concatenate([np.ones((len(inp), 1)), inp], axis=1)
    V = np.concatenate([np.ones((len(out_pts), 1)), out_pts], axis=1)
    beta = np.linalg.lstsq(Q, f, rcond=None)[0]
    lam = np.linalg.solve(Cm, f - Q @ beta)
    A = _c6(np.linalg.norm(out_pts[:, None, :] - inp[None, :, :], axis=2), support)
    return A @ lam + V @ beta


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import bench
    import oracle
    from precice_b200 import _lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, vpc = 2000, 25  # every cluster has more vertices than polynomial terms (the numpy fit has no rank-deficient branch)
    support = 2.0 * (3.0 * vpc / (4.0 * np.pi * n)) ** (1.0 / 3.0)
    inp = bench.halton_numpy(n, (2, 3, 5), 1)
    out = bench.halton_numpy(n, (7, 11, 13), 1)
    vals = np.sin(2 * np.pi * inp[:, 0]) * np.cos(2 * np.pi * inp[:, 1]) + inp[:, 2]
    m = oracle.PumMapping("consistent", 3, bench.BASIS, support, "separate", vpc, bench.OVERLAP, True)
    m.compute_mapping(inp, out)
    serial = m.map(vals, 1)
    radius = m.cluster_radius
    clusters = [m.cluster(c) for c in range(m.num_clusters)]
    centers = np.array([c["center"] for c in clusters])
    axis, lo_box, edge = 0, inp[:, 0].min(), inp[:, 0].max() - inp[:, 0].min()
    width = edge / N_BINS
    hist = np.bincount(np.clip(np.floor((centers[:, axis] - lo_box) / width), 0, N_BINS - 1).astype(int), minlength=N_BINS).astype(np.int64)
    L = _lib.lib()  # host-side helper of the product (no device needed)

    def slab(halo_bins):
        cuts = (C.c_int * (world + 1))()
        assert L.rbfmap_partition_slabs(hist.ctypes.data_as(C.POINTER(C.c_int64)), N_BINS, world, halo_bins, cuts) == 0
        lo = -np.inf if rank == 0 else lo_box + cuts[rank] * width
        hi = np.inf if rank == world - 1 else lo_box + cuts[rank + 1] * width
        return lo, hi

    results = {}
    # duplicate mode: owned output vertices, every cluster that reaches into the slab, normalised weights, no collective in map
    lo, hi = slab(int(np.ceil(radius / width)) + 1)
    own = (out[:, axis] >= lo) & (out[:, axis] < hi)
    part, processed = np.zeros(n), 0
    for c in clusters:
        if not (c["center"][axis] + radius > lo and c["center"][axis] - radius < hi):
            continue
        processed += 1
        sel = own[c["out_ids"]]
        if sel.any():
            s = _cluster_interpolant(inp[c["in_ids"]], out[c["out_ids"]][sel], vals[c["in_ids"]], support)
            part[c["out_ids"][sel]] += c["weights"][sel] * s
    assert np.all(part[~own] == 0.0)
    owned_exact = float(np.linalg.norm(part[own] - serial[own]) / np.linalg.norm(serial[own]))
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)  # only to assemble the complete field for the comparison
    results["duplicate"] = (float(np.linalg.norm(t.numpy() - serial) / np.linalg.norm(serial)), owned_exact, processed, int(own.sum()))

    # exchange mode: clusters owned by centre, weight sums and partial blends all-reduced
    lo, hi = slab(0)
    mine = [c for c in clusters if lo <= c["center"][axis] < hi]
    wsum, blend = np.zeros(n), np.zeros(n)
    for c in mine:
        w = _c2(np.linalg.norm(out[c["out_ids"]] - c["center"], axis=1) / radius)
        wsum[c["out_ids"]] += w
    tw = torch.from_numpy(wsum)
    dist.all_reduce(tw)  # computeMapping: the weight sums of an output vertex span clusters of two ranks at a cut
    for c in mine:
        w = _c2(np.linalg.norm(out[c["out_ids"]] - c["center"], axis=1) / radius) / wsum[c["out_ids"]]
        blend[c["out_ids"]] += w * _cluster_interpolant(inp[c["in_ids"]], out[c["out_ids"]], vals[c["in_ids"]], support)
    tb = torch.from_numpy(blend)
    dist.all_reduce(tb)  # map: partial blends
    results["exchange"] = (float(np.linalg.norm(blend - serial) / np.linalg.norm(serial)), len(mine))
    t2 = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t2, op=dist.ReduceOp.MAX)  # max over ranks, as bench.py does for the step time
    q.put((rank, results, len(clusters), float(t2[0])))
    dist.destroy_process_group()


def test_two_ranks_shard_one_mapping():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in results] == [0, 1]
    n_clusters = results[0][2]
    assert results[0][3] == results[1][3] == 11.0  # MAX over ranks
    for rank, res, _, _ in results:
        err, owned_err, processed, owned = res["duplicate"]
        assert err <= 1e-10 and owned_err <= 1e-10, (rank, err, owned_err)
        assert 0 < processed < n_clusters and owned > 0
        err, mine = res["exchange"]
        assert err <= 1e-10, (rank, err)
        assert 0 < mine < n_clusters
    # exchange mode: every cluster has exactly one owner; duplicate mode: the slabs partition the output mesh and the
    # boundary clusters are processed twice
    assert results[0][1]["exchange"][1] + results[1][1]["exchange"][1] == n_clusters
    assert results[0][1]["duplicate"][3] + results[1][1]["duplicate"][3] == 2000
    assert results[0][1]["duplicate"][2] + results[1][1]["duplicate"][2] > n_clusters
    # balanced within a few clusters (the cuts minimise the maximum)
    assert abs(results[0][1]["exchange"][1] - results[1][1]["exchange"][1]) <= 0.1 * n_clusters


def test_halton_generators_agree():
    import bench
    a = bench.halton_numpy(1000, (2, 3, 5), 1)
    b = bench.halton_numpy(1000, (2, 3, 5), 1001)
    c = bench.halton_numpy(2000, (2, 3, 5), 1)
    assert np.array_equal(np.concatenate([a, b]), c)
    assert a.min() > 0 and c.max() < 1
    import refcases
    assert np.allclose(refcases.halton(1000, (2, 3, 5)), a)


def test_reference_arm_prints_contract_line():
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-sample", "16000"], capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    assert d["impl"] == "reference" and d["unit"] == "vertices/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
