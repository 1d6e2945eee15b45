"""How often does rbf-pum-direct stop with "could not be assigned to any cluster" on the moving-surface workload of
remesh_bench.py? (The reference algorithm itself does on steep surfaces: oracle checked for individual cases.) Counts the
failing (rank, step) combinations per (amplitude, relative-overlap) on one GPU; used to choose the workload parameters."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import precice_b200 as pb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dev = torch.device("cuda", 0)
uin, uout = bench.halton_torch(n, (2, 3, 5), 1, dev), bench.halton_torch(n, (7, 11, 13), 1, dev)


def surface(uv, t, rank, amp):
    h = 1.0 / math.sqrt(n)
    x = uv[:, 0] + 0.3 * h * math.sin(0.7 * t) * torch.cos(2 * math.pi * uv[:, 1])
    y = uv[:, 1] + 0.3 * h * math.cos(0.9 * t) * torch.sin(2 * math.pi * uv[:, 0])
    z = 0.5 + (amp + 0.02 * math.sin(0.5 * t)) * torch.sin(2 * math.pi * x + 0.7 * rank) * torch.cos(2 * math.pi * y - 0.4 * rank)
    return torch.stack([x, y, z], dim=1).contiguous()


for amp, overlap, vpc, project in ((0.15, 0.15, 50, False), (0.15, 0.15, 50, True), (0.15, 0.3, 50, True), (0.15, 0.3, 50, False), (0.05, 0.15, 50, False), (0.15, 0.15, 100, False)):
    sup = 2.0 * math.sqrt(1.2 * vpc / (math.pi * n))
    m = pb.PartitionOfUnityMapping("consistent", 3, bench.BASIS, sup, "separate", vpc, overlap, project)
    fails, tot, ncl = 0, 0, 0
    for rank in range(8):
        for s in range(9):
            m.clear()
            tot += 1
            a, b = surface(uin, float(s), rank, amp), surface(uout, float(s), rank, amp)
            torch.cuda.synchronize()  # the library works on its own stream: the inputs must be complete before the call
            try:
                m.computeMapping(a, b)
                ncl = m.stats()["n_clusters"]
            except pb.MappingError as e:
                fails += 1
                if fails == 1:
                    print(f"  first failure (rank {rank}, step {s}): {e.status}: {str(e)[:160]}", flush=True)
    print(f"amplitude {amp} relative-overlap {overlap} vertices-per-cluster {vpc} project-to-input {project}: {fails} of {tot} steps fail; clusters {ncl}", flush=True)
