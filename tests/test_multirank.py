"""`-m "not gpu"`: the N > 1 path of bench.py — one process per rank, each rank maps its own partition (no data-path
collective), timing reduced with MAX over ranks — exercised with world_size 2 on the gloo backend."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import bench
    import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 3000
    # the partition of rank r: Halton indices [1 + r n, 1 + (r + 1) n), as in bench.py
    inp = bench.halton_numpy(n, (2, 3, 5), 1 + rank * n)
    out = bench.halton_numpy(n, (7, 11, 13), 1 + rank * n)
    vals = np.sin(2 * np.pi * inp[:, 0]) * np.cos(2 * np.pi * inp[:, 1]) + inp[:, 2]
    m = oracle.PumMapping("consistent", 3, bench.BASIS, bench.support_radius(n), "separate", bench.VPC, bench.OVERLAP, True)
    m.compute_mapping(inp, out)
    res = m.map(vals, 1)
    t = torch.tensor([10.0 + rank, float(res.sum())], dtype=torch.float64)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)  # max over ranks, as bench.py does for the step time
    gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, t)
    q.put((rank, float(mx[0]), [g.tolist() for g in gathered], inp[0].tolist()))
    dist.destroy_process_group()


def test_two_ranks_map_disjoint_partitions():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in results] == [0, 1]
    assert results[0][1] == results[1][1] == 11.0  # MAX over ranks
    assert results[0][2] == results[1][2]  # both ranks see both partial results
    assert results[0][3] != results[1][3]  # the ranks mapped different partitions
    sums = [g[1] for g in results[0][2]]
    assert all(np.isfinite(sums)) and sums[0] != sums[1]


def test_halton_partitions_are_disjoint_and_deterministic():
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
