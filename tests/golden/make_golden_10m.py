"""Generates tests/golden/pum_cfg4_10m.npz: the CPU oracle's mapped values on the REAL headline workload (BASELINE.json
configs[3]: rbf-pum-direct, compact-polynomial-c6, 3-D unit cube, 10M -> 10M Halton vertices, vertices-per-cluster 50,
relative-overlap 0.15, project-to-input, polynomial separate, support radius = 2 x the nominal cluster radius) at 10 000
sampled vertices of the result mesh, for consistent and conservative constraints, scalar and 3-vector data, together with
the cluster count and the cluster radius.

bench.py asserts its result against this fixture in every run at the default size (the oracle cannot map 10M vertices
inside the bench; oracle.compute_mapping_for_samples builds only the ~46 000 clusters that hold a sampled vertex, with
the clustering and the normalised weights of the complete mapping, and is bit-identical to the full oracle at the
samples - tests/test_oracle_golden.py::test_sampled_oracle_matches_full_oracle).

    python tests/golden/make_golden_10m.py [--vertices 10000000]     # a few minutes and ~12 GB of host memory
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import bench  # the workload definition lives there (halton_numpy, support_radius, BASIS, VPC, OVERLAP)
import oracle as orc

N_SAMPLES = 10_000


def field(xyz, dims):
    f = np.sin(2 * np.pi * xyz[:, 0]) * np.cos(2 * np.pi * xyz[:, 1]) + xyz[:, 2]
    return f if dims == 1 else np.stack([f, 2 * f + 1, -f], axis=1).reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vertices", dest="n", type=int, default=10_000_000)
    ap.add_argument("--out", default=os.path.join(HERE, "pum_cfg4_10m.npz"))
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    n = a.n
    t0 = time.time()
    inp = bench.halton_numpy(n, (2, 3, 5), 1)
    out = bench.halton_numpy(n, (7, 11, 13), 1)
    support = bench.support_radius(n)
    samples = np.sort(np.random.default_rng(20261020).choice(n, N_SAMPLES, replace=False)).astype(np.int32)
    data = {"n": n, "samples": samples, "support_radius": support}
    print(f"meshes {time.time() - t0:.1f} s", flush=True)
    for constraint in ("consistent", "conservative"):
        m = orc.PumMapping(constraint, 3, bench.BASIS, support, "separate", bench.VPC, bench.OVERLAP, True)
        t0 = time.time()
        m.compute_mapping_for_samples(inp, out, samples, threads=a.threads)
        print(f"{constraint}: computeMapping {time.time() - t0:.1f} s, {m.num_clusters} clusters, radius {m.cluster_radius!r}", flush=True)
        data[f"{constraint}_n_clusters"] = m.num_clusters
        data[f"{constraint}_radius"] = m.cluster_radius
        for dims in (1, 3):
            res = m.map(field(inp, dims), dims, threads=1)  # single thread: the reference's summation order
            data[f"{constraint}_{dims}"] = res.reshape(n, dims)[samples].copy()
        del m
    np.savez_compressed(a.out, **data)
    print("wrote", a.out, os.path.getsize(a.out), "bytes")


if __name__ == "__main__":
    main()
