#!/usr/bin/env python
"""a small run of the round-2 kernels for `compute-sanitizer --tool racecheck`: kernel 1t + 2v + 3v (graph-row route forced on
a Cora-shaped graph) and kernel S (three classes, extended)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tlc-gnn_b200"))
import bench
from tlc_b200 import _lib as L, api
c, labels, ne, csr, perm = bench.make_workload("cora")
g = api.VicinityGraph(*csr, device=0)
tg = bench.batch_targets(ne, perm, 0, 0, 1, 192)
g.vicinity_pi(tg, hop=2, flags=L.F_NORM | L.F_NO_SMALL | L.F_DIRECT)
print("table route rows:", g.last_table() if hasattr(g, "last_table") else "?")
g.vicinity_pi(tg, hop=2, flags=L.F_NORM | L.F_EXTENDED)
print("kernel S:", g.last_small())
g.close()
