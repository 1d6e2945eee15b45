"""Scratch diagnostics run on the GPU box (not part of the test-suite)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle, refcases
import precice_b200 as pb

print("files:", sorted(os.listdir(ROOT)))
pts3 = np.array([(float(i), float(j), float(k)) for i in range(10) for j in range(i, 10) for k in range(i, 10)])
for project in (False, True):
    g = pb.PartitionOfUnityMapping("consistent", 3, "compact-polynomial-c6", 10.0, "separate", 10, 0.15, project)
    o = oracle.PumMapping("consistent", 3, "compact-polynomial-c6", 10.0, "separate", 10, 0.15, project)
    o.compute_mapping(pts3, pts3)
    try:
        g.compute_mapping(pts3, pts3)
    except pb.MappingError as e:
        print("ERR", e.status, str(e)[:80])
    st = g.stats()
    print("project", project, "gpu radius", st["cluster_radius"], "n", st["n_clusters"], "| oracle", o.cluster_radius, o.num_clusters)
    nc = st["n_clusters"]
    centers = np.zeros((nc, 3)); io = np.zeros(nc + 1, dtype=np.int64); oo = np.zeros(nc + 1, dtype=np.int64)
    g._check(g._L.rbfmap_pum_get_clusters(g._h, centers.ctypes.data, None, None))
    oc = np.array([o.cluster(c)["center"] for c in range(o.num_clusters)])
    print(" gpu centers[:5]", centers[:5].tolist()); print(" orc centers[:5]", oc[:5].tolist())
    sg = {tuple(np.round(c, 9)) for c in centers}; so = {tuple(np.round(c, 9)) for c in oc}
    print(" only gpu:", sorted(sg - so)[:10]); print(" only oracle:", sorted(so - sg)[:10])
