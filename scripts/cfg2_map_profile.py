"""BASELINE configs[1] (rbf-global-direct, compact-polynomial-c2, 50k -> 50k): one computeMapping + a few maps; run under ncu for
the launch list of the map. Not timed: numbers printed here are never bench values."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, precice_b200 as pb
dev = torch.device("cuda", 0); n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
inp = bench.halton_torch(n, (2, 3, 5), 1, dev); out = bench.halton_torch(n, (7, 11, 13), 1, dev); vals = bench.field_torch(inp).contiguous()
m = pb.RadialBasisFctMapping("consistent", 3, "compact-polynomial-c2", 0.2, polynomial="separate")
m.computeMapping(inp, out)
for _ in range(3):
    r = m.map(vals, 1)
    print("map ms", round(m.stats()["ms_map_kernels"], 3), flush=True)
