"""One or more device-resident steps (clear + computeMapping + map) of the bench workload; run under ncu or with
RBFMAP_TRACE=1. Not timed: numbers printed here are never bench values."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import precice_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--vertices", dest="n", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--constraint", default="consistent")
ap.add_argument("--dims", type=int, default=1)
ap.add_argument("--virtual-rank", default=None, help="R/W: this GPU plays rank R of a mapping sharded over W ranks (group without communicator)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
inp = bench.halton_torch(a.n, (2, 3, 5), 1, dev)
out = bench.halton_torch(a.n, (7, 11, 13), 1, dev)
vals = bench.field_torch(inp if a.constraint == "consistent" else out).contiguous()
if a.dims > 1:
    vals = torch.stack([vals * (k + 1) for k in range(a.dims)], dim=1).contiguous().reshape(-1)
m = pb.PartitionOfUnityMapping(a.constraint, 3, bench.BASIS, bench.support_radius(a.n), "separate", bench.VPC, bench.OVERLAP, True)
if a.virtual_rank:
    r_, w_ = (int(x) for x in a.virtual_rank.split("/"))
    m.distributedInit(None, r_, w_)
for s in range(a.steps):
    if m.hasComputedMapping():
        m.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.computeMapping(inp, out)
    t1 = time.perf_counter()
    r = m.map(vals, a.dims)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = m.stats()
    print(f"step {s}: compute {1e3*(t1-t0):.2f} ms (events {st['ms_compute_kernels']:.2f}, factor {st['ms_assembly_factor']:.2f}) "
          f"map {1e3*(t2-t1):.2f} ms (events {st['ms_map_kernels']:.2f}) clusters {st['n_clusters']} launches {st['kernel_launches']}",
          flush=True)
