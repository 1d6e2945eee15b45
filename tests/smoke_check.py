"""Body of __graft_entry__.smoke(): one small rbf-pum-direct computeMapping + map (consistent and conservative)
on cuda:0 through the C ABI, checked against the CPU oracle. Lives under tests/ because it uses the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run() -> None:
    import oracle
    import precice_b200 as pb
    import refcases
    if pb.device_count() == 0:
        raise RuntimeError("smoke(): no CUDA device; precice_b200 has no CPU fallback")
    n = 4000
    inp, out = refcases.halton(n, (2, 3, 5)), refcases.halton(n, (7, 11, 13))
    vals = refcases.field(inp)
    for constraint in ("consistent", "conservative"):
        g = pb.PartitionOfUnityMapping(constraint, 3, "compact-polynomial-c6", 0.3, "separate", 50, 0.15, True)
        o = oracle.PumMapping(constraint, 3, "compact-polynomial-c6", 0.3, "separate", 50, 0.15, True)
        g.computeMapping(inp, out)
        o.compute_mapping(inp, out)
        rg, ro = g.map(vals, 1), o.map(vals, 1)
        err = refcases.rel_l2(rg, ro)
        st = g.stats()
        print(f"smoke {constraint}: clusters={st['n_clusters']} radius={st['cluster_radius']:.6f} "
              f"rel_l2_vs_oracle={err:.3e} launches={st['kernel_launches']}")
        assert st["n_clusters"] == o.num_clusters
        assert err <= 1e-10, err
    print("smoke OK")


if __name__ == "__main__":
    run()
