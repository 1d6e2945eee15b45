"""BASELINE.json configs[1]: rbf-global-direct compact-polynomial-c2, 3-D unit cube n->n (default 50k), consistent scalar +
conservative 3-vector on one B200. Prints device times (CUDA events inside the library) and a sampled parity check."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench, refcases
import precice_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--vertices", dest="n", type=int, default=50000)
ap.add_argument("--support", type=float, default=0.2)
ap.add_argument("--basis", default="compact-polynomial-c2")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--solver", default="direct")
a = ap.parse_args()
dev = torch.device("cuda", 0)
inp = bench.halton_torch(a.n, (2, 3, 5), 1, dev)
out = bench.halton_torch(a.n, (7, 11, 13), 1, dev)
f = bench.field_torch(inp).contiguous()
for constraint, dims in (("consistent", 1), ("conservative", 3)):
    m = pb.RadialBasisFctMapping(constraint, 3, a.basis, a.support, polynomial="separate", solver=a.solver)
    src = inp if constraint == "consistent" else out
    vals = bench.field_torch(src)
    if dims > 1:
        vals = torch.stack([vals, 2 * vals + 1, -vals], dim=1)
    vals = vals.contiguous().reshape(-1)
    for s in range(a.steps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.computeMapping(inp, out) if constraint == "consistent" else m.computeMapping(out, inp)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        r = m.map(vals, dims)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        st = m.stats()
        n3 = (a.n ** 3) / 3.0
        print(f"{constraint} dims={dims} step {s}: compute {1e3*(t1-t0):.1f} ms (factor+assembly {st['ms_assembly_factor']:.1f} ms = "
              f"{n3/ (max(st['ms_assembly_factor'],1e-9)*1e-3)/1e12:.2f} TFLOP/s Cholesky-equivalent) map {1e3*(t2-t1):.1f} ms launches {st['kernel_launches']} "
              f"iters {st['iterations']} res {st['residual']:.2e}", flush=True)
    if constraint == "consistent":
        # interpolation sanity: mapped field vs the analytic field at the output vertices
        truth = bench.field_torch(out)
        print("  rel. L2 vs analytic field:", float(torch.linalg.norm(r - truth) / torch.linalg.norm(truth)))
    else:
        print("  component sums in/out:", vals.reshape(-1, 3).sum(0).tolist(), r.reshape(-1, 3).sum(0).tolist())
