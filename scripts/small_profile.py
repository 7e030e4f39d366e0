#!/usr/bin/env python
"""per-stage clock ticks of kernel S (needs a -DSMALL_PROFILE build selected with TLC_LIB): one line per size class
   python scripts/small_profile.py <workload> <hop> <extended> <batch>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tlc-gnn_b200"))
import numpy as np
import bench
from tlc_b200 import _lib as L, api
name, hop, ext, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
neg = 1 if name == "collab" else 0
c, labels, ne, csr, perm = bench.make_workload(name, "edge", neg)
g = api.VicinityGraph(*csr, device=0)
fl = L.F_NORM | (L.F_EXTENDED if ext else 0)
g.vicinity_pi(bench.batch_targets(ne, perm, 0, 0, 1, B), hop=hop, flags=fl)   # warm-up (ball cache)
os.environ["SMALL_PROFILE_DUMP"] = "1"
g.vicinity_pi(bench.batch_targets(ne, perm, 1, 0, 1, B), hop=hop, flags=fl)   # prints + resets the warm-up's ticks
print("---- %s hop %d ext %d batch %d: the step below ----" % (name, hop, ext, B), file=sys.stderr, flush=True)
g.vicinity_pi(bench.batch_targets(ne, perm, 2, 0, 1, B), hop=hop, flags=fl)   # prints step 1's ticks
print(g.last_small(), file=sys.stderr)
g.close()
