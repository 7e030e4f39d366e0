#!/usr/bin/env python
"""per-phase clock ticks of kernel 1t (needs the -DT1_PROFILE build, selected with TLC_LIB):
   TLC_LIB=tlc-gnn_b200/tlc_b200/libtlc_b200_prof.so python scripts/table_profile.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tlc-gnn_b200"))
import bench
from tlc_b200 import _lib as L, api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c, labels, ne, csr, perm = bench.make_workload("computers")
g = api.VicinityGraph(*csr, device=0)
fl = L.F_NORM
g.vicinity_pi(bench.batch_targets(ne, perm, 0, 0, 1, B), hop=2, flags=fl)   # warm-up (ball cache, tables)
os.environ["T1_PROFILE_DUMP"] = "1"
for s in range(1, 3):
    print("---- step %d ----" % s, file=sys.stderr, flush=True)
    g.vicinity_pi(bench.batch_targets(ne, perm, s, 0, 1, B), hop=2, flags=fl)
g.close()
