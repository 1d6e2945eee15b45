"""GPU-vs-oracle deviation of the reference harness case (2-D, 2 components, conservative): prints error statistics."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refcases, oracle as orc
import precice_b200 as pb
for constraint in ("conservative", "consistent"):
    for dim, ncomp in ((2, 2), (3, 3)):
        inp, out = refcases.reference_testing_meshes(dim)
        src = inp if constraint == "consistent" else out
        vals = refcases.reference_testing_values(src.shape[0], ncomp)
        a, b = (inp, out) if constraint == "consistent" else (out, inp)
        g = pb.PartitionOfUnityMapping(constraint, dim, "compact-polynomial-c6", 3.0, "off", 25, 0.15, True)
        o = orc.PumMapping(constraint, dim, "compact-polynomial-c6", 3.0, "off", 25, 0.15, True)
        g.compute_mapping(a, b); o.compute_mapping(a, b)
        rg, ro = g.map(vals, ncomp), o.map(vals, ncomp)
        d = np.abs(rg - ro)
        print(constraint, dim, ncomp, "relL2 %.3e" % refcases.rel_l2(rg, ro), "max|d| %.3e" % d.max(), "max|ro| %.3e" % np.abs(ro).max(),
              "max d/max|ro| %.3e" % (d.max() / np.abs(ro).max()), "n", a.shape[0], b.shape[0], flush=True)
