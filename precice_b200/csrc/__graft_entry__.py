"""Driver entry points: build() compiles every CUDA source for sm_100a (and the CPU oracle used as checker);
smoke() runs one small rbf-pum-direct mapping on cuda:0 through the C ABI and checks it against the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build() -> None:
    # the product: precice_b200/librbfmap.so (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo, see csrc/Makefile)
    subprocess.run(["make", "-C", os.path.join(ROOT, "precice_b200", "csrc"), "-j", str(os.cpu_count() or 4)], check=True)
    # the checker: oracle/_build/librbf_oracle.so (g++). The reference itself cannot be compiled here
    # (needs Eigen/Boost/libxml2), so there is no oracle/_ref.
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    import precice_b200
    from precice_b200 import _lib
    L = _lib.lib()
    for name in _lib.EXPORTS:
        getattr(L, name)
    assert precice_b200.PartitionOfUnityMapping is not None


def smoke() -> None:
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
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
        print(f"smoke {constraint}: clusters={st['n_clusters']} radius={st['cluster_radius']:.6f} rel_l2_vs_oracle={err:.3e} "
              f"launches={st['kernel_launches']}")
        assert st["n_clusters"] == o.num_clusters
        assert err <= 1e-10, err
    print("smoke OK")


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
