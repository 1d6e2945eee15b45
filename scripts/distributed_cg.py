"""rbf-global-iterative sharded over N GPUs (one process per GPU, launched with torchrun): rows of the matrix-free CG
operator are split across the ranks, NCCL all-gathers the direction vector and all-reduces the dot products.
Checks the sharded result against the single-GPU result of rank 0 and prints device times (max over ranks).
BASELINE.json configs[2]: gaussian, 3-D, 500k -> 500k."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import bench
import precice_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--vertices", dest="n", type=int, default=500000)
ap.add_argument("--basis", default="gaussian")
ap.add_argument("--param", type=float, default=91.05)
ap.add_argument("--rtol", type=float, default=1e-9)
ap.add_argument("--dims", type=int, default=1)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
inp = bench.halton_torch(a.n, (2, 3, 5), 1, dev)
out = bench.halton_torch(a.n, (7, 11, 13), 1, dev)
vals = bench.field_torch(inp)
if a.dims > 1:
    vals = torch.stack([vals * (k + 1) for k in range(a.dims)], dim=1)
vals = vals.contiguous().reshape(-1)

def run(sharded):
    m = pb.RadialBasisFctMapping("consistent", 3, a.basis, a.param, polynomial="separate", solver="iterative", solver_rtol=a.rtol,
                                 device_id=local)
    if sharded and world > 1:
        ident = [pb.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        m.distributedInit(ident[0], rank, world)
    m.computeMapping(inp, out)
    torch.cuda.synchronize()
    if sharded and world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    r = m.map(vals, a.dims)  # first call: includes NCCL channel set-up when sharded
    if a.dims >= 1:
        m.clear()
        m.computeMapping(inp, out)  # resets the warm start, so that the second call solves from zero again
        torch.cuda.synchronize()
        if sharded and world > 1:
            dist.barrier()
    t0 = time.perf_counter()
    r = m.map(vals, a.dims)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    st = m.stats()
    del m
    return r, t, st

r_sh, t_sh, st = run(True)
tt = torch.tensor([t_sh], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    r_1, t_1, st1 = run(False)  # single-GPU reference on rank 0
    err = float(torch.linalg.norm(r_sh - r_1) / torch.linalg.norm(r_1))
    print(f"world={world} n={a.n} dims={a.dims}: sharded map {1e3*float(tt):.1f} ms ({st['iterations']} it, res {st['residual']:.2e}) | "
          f"single-GPU map {1e3*t_1:.1f} ms ({st1['iterations']} it) | rel.L2 sharded vs single {err:.2e}", flush=True)
    assert err <= 100 * a.rtol, err
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
