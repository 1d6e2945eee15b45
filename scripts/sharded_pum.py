"""ONE rbf-pum-direct mapping sharded over N GPUs (one process per GPU, launched with torchrun).

For each case (shard mode x constraint x data dimension x host/device pointers) every rank maps the same replicated meshes
through a sharded handle and through an ordinary single-device handle and compares; rank 0 optionally compares with the
CPU oracle as well (small meshes). Asserts 1e-10 relative L2 and prints `SHARDED PUM OK` on success.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/sharded_pum.py --vertices 20000 --oracle
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import bench
import precice_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--vertices", dest="n", type=int, default=20000)
ap.add_argument("--vpc", type=int, default=50)
ap.add_argument("--oracle", action="store_true", help="rank 0 also compares with the CPU oracle")
ap.add_argument("--time", action="store_true", help="print step times of the sharded mapping (device pointers)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
TOL = 1e-10


def rel_l2(x, y):
    x, y = np.asarray(x), np.asarray(y)
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))


n = a.n
support = 2.0 * (3.0 * a.vpc / (4.0 * np.pi * n)) ** (1.0 / 3.0)
inp = bench.halton_numpy(n, (2, 3, 5), 1)
out = bench.halton_numpy(n, (7, 11, 13), 1)


def values(constraint, dims):
    x = inp  # Mapping::input() carries the data in both directions
    f = np.sin(2 * np.pi * x[:, 0]) * np.cos(2 * np.pi * x[:, 1]) + x[:, 2]
    return f if dims == 1 else np.stack([f, 2 * f + 1, -f], axis=1).reshape(-1)


def make(constraint, shard_mode=None):
    m = pb.PartitionOfUnityMapping(constraint, 3, "compact-polynomial-c6", support, "separate", a.vpc, 0.15, True, device_id=local,
                                   shard_mode=shard_mode or "duplicate")
    if shard_mode is not None:
        ident = [pb.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        m.distributedInit(ident[0], rank, world)
    return m


failures = []
for constraint in ("consistent", "conservative"):
    for dims in (1, 3):
        vals = values(constraint, dims)
        single = make(constraint)
        single.computeMapping(inp, out)
        ref = single.map(vals, dims)
        oracle_res = None
        if a.oracle and rank == 0:
            import oracle
            o = oracle.PumMapping(constraint, 3, "compact-polynomial-c6", support, "separate", a.vpc, 0.15, True)
            o.compute_mapping(inp, out)
            oracle_res = o.map(vals, dims)
            e = rel_l2(ref, oracle_res)
            print(f"[single vs oracle] {constraint} dims={dims}: {e:.2e}", flush=True)
        for mode in ("duplicate", "exchange"):
            for where in ("host", "device"):
                m = make(constraint, mode)
                if where == "host":
                    m.computeMapping(inp, out)
                    res = m.map(vals, dims)
                    # rbfmap_map_sliced: every rank reads back its slice of the complete result; the slices tile the result
                    part, b, e = m.mapSliced(vals, dims, out=np.full(ref.size, np.nan))
                    untouched = np.ones(ref.size, bool)
                    untouched[b * dims:e * dims] = False
                    if not np.all(np.isnan(part[untouched])):
                        failures.append(f"{constraint}/{mode}/dims={dims}: mapSliced wrote outside its slice")
                    pieces = [None] * world
                    dist.all_gather_object(pieces, (b, e, part[b * dims:e * dims]))
                    tiled = np.concatenate([p[2] for p in sorted(pieces, key=lambda p: p[0])])
                    es = rel_l2(tiled, ref) if tiled.size == ref.size else float("inf")
                    if es > TOL:
                        failures.append(f"{constraint}/{mode}/dims={dims}: mapSliced differs from the single-device result by {es:.2e}")
                else:
                    ti, to = torch.from_numpy(inp).to(dev), torch.from_numpy(out).to(dev)
                    tv = torch.from_numpy(np.ascontiguousarray(vals)).to(dev)
                    m.computeMapping(ti, to)
                    res = m.map(tv, dims).cpu().numpy()
                st = m.stats()
                exchange = mode == "exchange" or constraint == "conservative"
                if exchange:
                    full = res  # every rank holds the complete result
                else:
                    t = torch.from_numpy(res).to(dev)  # the complete result is the sum of the owned parts
                    dist.all_reduce(t)
                    full = t.cpu().numpy()
                    own = (out[:, st["shard_axis"]] >= st["shard_lo"]) & (out[:, st["shard_axis"]] < st["shard_hi"])
                    if not np.all(res.reshape(n, dims)[~own] == 0.0):
                        failures.append(f"{constraint}/{mode}/{where}/dims={dims}: wrote outside the owned slab")
                e = rel_l2(full, ref)
                line = (f"[rank {rank}] {constraint:12s} dims={dims} {mode:9s} {where:6s}: clusters {st['n_clusters']}/{st['n_clusters_global']} "
                        f"slab axis {st['shard_axis']} [{st['shard_lo']:.4f}, {st['shard_hi']:.4f})  rel.L2 vs single {e:.2e}")
                if oracle_res is not None:
                    eo = rel_l2(full, oracle_res)
                    line += f"  vs oracle {eo:.2e}"
                    if eo > TOL:
                        failures.append(line)
                print(line, flush=True)
                if e > TOL:
                    failures.append(line)
                if constraint == "conservative":  # component sums are conserved
                    s_in, s_out = vals.reshape(n, dims).sum(0), full.reshape(n, dims).sum(0)
                    if np.max(np.abs(s_in - s_out) / np.maximum(np.abs(s_in), 1.0)) > 1e-9:
                        failures.append(f"{constraint}/{mode}/{where}/dims={dims}: sums not conserved {s_in} {s_out}")
                del m
        del single

# Errors are agreed on by all ranks (one grouped all-reduce at the end of computeMapping): a stray vertex that only one
# rank owns, or a singular cluster that only one rank factorises, raises the SAME error on every rank - none is left
# waiting in a collective - and the singular matrix wins over the unassigned vertex like in the reference's order.
stray = np.concatenate([out, [[0.5, 0.5, 3.0]]])
dup = inp.copy()
dup[1] = dup[0]  # two identical in-vertices: one cluster matrix is singular
for mode in ("duplicate", "exchange"):
    for label, ci, co, want in (("unassigned", inp, stray, "UNASSIGNED"), ("singular", dup, out, "SINGULAR"), ("both", dup, stray, "SINGULAR")):
        m = make("consistent", mode)
        got = "none"
        try:
            m.computeMapping(ci, co)
        except pb.MappingError as e:
            got = e.status
        print(f"[rank {rank}] error agreement {mode:9s} {label:10s}: {got}", flush=True)
        if got != want:
            failures.append(f"error agreement {mode}/{label}: rank {rank} got {got}, expected {want}")
        del m

if a.time:
    ti, to = torch.from_numpy(inp).to(dev), torch.from_numpy(out).to(dev)
    tv = torch.from_numpy(values("consistent", 1)).to(dev)
    for mode in ("duplicate", "exchange"):
        m = make("consistent", mode)
        ts = []
        for it in range(6):
            if m.hasComputedMapping():
                m.clear()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            m.computeMapping(ti, to)
            m.map(tv, 1)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        st = m.stats()
        t = torch.tensor([min(ts[2:])], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"[time] {mode}: {t.item() * 1e3:.2f} ms/step (max over ranks), replicated {st['ms_replicated']:.2f} ms, "
                  f"map {st['ms_map_kernels']:.2f} ms (collectives {st['ms_collectives']:.2f} ms)", flush=True)
        del m

bad = torch.tensor([len(failures)], device=dev)
dist.all_reduce(bad)
if rank == 0:
    for f in failures:
        print("FAIL", f, flush=True)
    print("SHARDED PUM OK" if bad.item() == 0 else f"SHARDED PUM FAILED ({bad.item()})", flush=True)
dist.destroy_process_group()
sys.exit(0 if bad.item() == 0 else 1)
