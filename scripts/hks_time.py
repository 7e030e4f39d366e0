#!/usr/bin/env python
"""wall time of the heat-kernel-signature filtration on 2048 PubMed-shaped node vicinities (2-hop, KD flags)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tlc-gnn_b200"))
import numpy as np
import bench
from tlc_b200 import _lib as L, api
c, labels, ne, csr, perm = bench.make_workload("pubmed")
g = api.VicinityGraph(*csr, device=0)
ids = np.arange(2048, dtype=np.int32)
nodes = np.stack([ids, ids], 1)
kd = L.F_NORM | L.F_EXTENDED | L.F_KEEP_ZERO | L.F_NORM_EPS
g.set_hks_time(0.1)
for flt, name in ((L.F_FILT_DEGREE, "degree"), (L.F_FILT_HKS, "hks")):
    g.vicinity_pi(nodes[:64], hop=2, mode=L.MODE_NODE, flags=kd | flt)
    t0 = time.perf_counter()
    g.vicinity_pi(nodes, hop=2, mode=L.MODE_NODE, flags=kd | flt)
    print("%s filtration, 2048 node vicinities, whole call: %.1f ms" % (name, (time.perf_counter() - t0) * 1e3), flush=True)
g.close()
