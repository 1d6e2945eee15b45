"""Just-in-time mapping throughput (not part of bench.py's headline metric): mesh of N Halton vertices, read mapping
(updateMappingDataCache + mapConsistentAt at M points) and write mapping (mapConservativeAt + completeJustInTimeMapping),
host pointers (H2D / D2H included), device vs the CPU oracle on a sample for parity."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, refcases
import precice_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--vertices", dest="n", type=int, default=2_000_000)
ap.add_argument("--points", dest="m", type=int, default=2_000_000)
ap.add_argument("--oracle-sample", type=int, default=20000)
a = ap.parse_args()
mesh = bench.halton_torch(a.n, (2, 3, 5), 1, "cpu").numpy()
pts = 0.02 + 0.96 * bench.halton_torch(a.m, (7, 11, 13), 1, "cpu").numpy()
vals = refcases.field(mesh)
support = bench.support_radius(a.n)
m = pb.PartitionOfUnityMapping("consistent", 3, bench.BASIS, support, "separate", bench.VPC, bench.OVERLAP, True)
for rep in range(3):
    if m.hasComputedMapping():
        m.clear()
    t0 = time.perf_counter(); m.computeMappingJustInTime(mesh); t1 = time.perf_counter()
    c = m.initializeMappingDataCache(1)
    t2 = time.perf_counter(); m.updateMappingDataCache(c, vals); t3 = time.perf_counter()
    r = m.mapConsistentAt(pts, c); t4 = time.perf_counter()
    print(f"read : computeMapping {1e3*(t1-t0):.1f} ms, updateCache {1e3*(t3-t2):.1f} ms, mapConsistentAt({a.m}) {1e3*(t4-t3):.1f} ms = {a.m/(t4-t3)/1e6:.1f} M points/s", flush=True)
w = pb.PartitionOfUnityMapping("conservative", 3, bench.BASIS, support, "separate", bench.VPC, bench.OVERLAP, True)
loads = refcases.field(pts) + 0.5
for rep in range(3):
    if w.hasComputedMapping():
        w.clear()
    w.computeMappingJustInTime(mesh)
    c = w.initializeMappingDataCache(1)
    out = np.zeros(a.n)
    t0 = time.perf_counter(); w.mapConservativeAt(pts, loads, c); t1 = time.perf_counter()
    w.completeJustInTimeMapping(c, out); t2 = time.perf_counter()
    print(f"write: mapConservativeAt({a.m}) {1e3*(t1-t0):.1f} ms = {a.m/(t1-t0)/1e6:.1f} M points/s, completeJustInTimeMapping {1e3*(t2-t1):.1f} ms, "
          f"sum {out.sum():.6f} vs {loads.sum():.6f}", flush=True)
# parity on a sample against the oracle
import oracle as orc
ns = a.oracle_sample
smesh = mesh[:ns] if False else bench.halton_torch(ns, (2, 3, 5), 1, "cpu").numpy()
spts = pts[: ns // 2]
sup = bench.support_radius(ns)
g = pb.PartitionOfUnityMapping("consistent", 3, bench.BASIS, sup, "separate", bench.VPC, bench.OVERLAP, True)
o = orc.PumMapping("consistent", 3, bench.BASIS, sup, "separate", bench.VPC, bench.OVERLAP, True)
g.computeMappingJustInTime(smesh); o.compute_mapping_jit(smesh, threads=os.cpu_count() or 1)
cg, co = g.initializeMappingDataCache(1), o.initialize_cache(1)
sv = refcases.field(smesh)
g.updateMappingDataCache(cg, sv); t0 = time.perf_counter(); o.update_cache(co, sv); t1 = time.perf_counter()
rg = g.mapConsistentAt(spts, cg); ro = o.map_consistent_at(spts, co); t2 = time.perf_counter()
print(f"parity (n={ns}): rel L2 {refcases.rel_l2(rg, ro):.3e}; oracle (1 thread) updateCache {1e3*(t1-t0):.1f} ms, mapConsistentAt({len(spts)}) {1e3*(t2-t1):.1f} ms "
      f"= {len(spts)/(t2-t1)/1e6:.3f} M points/s", flush=True)
