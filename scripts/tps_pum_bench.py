import sys
sys.path.insert(0, '/root/repo')
import torch, bench, precice_b200 as pb
dev = torch.device("cuda", 0); n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
inp = bench.halton_torch(n, (2, 3, 5), 1, dev); out = bench.halton_torch(n, (7, 11, 13), 1, dev); vals = bench.field_torch(inp).contiguous()
m = pb.PartitionOfUnityMapping("consistent", 3, "thin-plate-splines", 0.0, "separate", 50, 0.15, False)
for _ in range(3):
    if m.hasComputedMapping(): m.clear()
    m.computeMapping(inp, out); r = m.map(vals, 1); st = m.stats()
    print("tps 2M: compute", round(st["ms_compute_kernels"], 2), "inverse", round(st["ms_assembly_factor"], 2), "map", round(st["ms_map_kernels"], 2),
          "err", float(torch.linalg.norm(r - bench.field_torch(out)) / torch.linalg.norm(bench.field_torch(out))))
