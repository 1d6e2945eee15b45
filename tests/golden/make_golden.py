"""Generates tests/golden/pum_halton.npz: mapped fields of the CPU oracle (oracle/, pinned by the reference's own golden
literals, see tests/test_oracle_golden.py) on small Halton meshes. The `-m gpu` tests compare the CUDA path with these
committed vectors, the CPU tests check that the oracle still reproduces them.

    python tests/golden/make_golden.py        # rewrites pum_halton.npz
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refcases
import oracle as orc

CASES = [  # name, constraint, dim, basis, polynomial, n_in, n_out, dims
    ("cons3d_c6", "consistent", 3, "compact-polynomial-c6", "separate", 4000, 3500, 1),
    ("consv3d_c6", "conservative", 3, "compact-polynomial-c6", "separate", 4000, 3500, 1),
    ("cons2d_c2_vec", "consistent", 2, "compact-polynomial-c2", "separate", 3000, 3000, 2),
    ("cons3d_c4_off", "consistent", 3, "compact-polynomial-c4", "off", 3000, 2500, 3),
]


def inputs(dim, n_in, n_out, dims, constraint):
    inp = refcases.halton(n_in, (2, 3, 5)[:dim])
    out = refcases.halton(n_out, (7, 11, 13)[:dim])
    src = inp
    vals = np.stack([refcases.field(src) * (k + 1) + k for k in range(dims)], axis=1).reshape(-1)
    return inp, out, vals


def support_of(constraint, dim, basis, polynomial, inp, out):
    probe = orc.PumMapping(constraint, dim, basis, 1.0, polynomial, 50, 0.15, True)
    probe.compute_mapping(inp, out)
    return 2.0 * probe.cluster_radius  # support = 2 x cluster radius


def run_case(case):
    name, constraint, dim, basis, polynomial, n_in, n_out, dims = case
    inp, out, vals = inputs(dim, n_in, n_out, dims, constraint)
    support = support_of(constraint, dim, basis, polynomial, inp, out)
    m = orc.PumMapping(constraint, dim, basis, support, polynomial, 50, 0.15, True)
    m.compute_mapping(inp, out)
    return support, m.map(vals, dims), m.num_clusters, m.cluster_radius


if __name__ == "__main__":
    data = {}
    for case in CASES:
        support, res, nc, radius = run_case(case)
        data[case[0] + "/support"] = np.array(support)
        data[case[0] + "/result"] = res
        data[case[0] + "/clusters"] = np.array(nc)
        data[case[0] + "/radius"] = np.array(radius)
        print(case[0], "support", support, "clusters", nc, "|result|", np.linalg.norm(res))
    np.savez_compressed(os.path.join(HERE, "pum_halton.npz"), **data)
