"""diagnostic: where does the host-buffer call (tlc_vicinity_pi) spend its time? (run on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tlc-gnn_b200"))
sys.path.insert(0, ROOT)
import numpy as np
os.environ["TLC_STAGE_TIMING"] = "1"
import bench
from tlc_b200 import api, _lib as L

class A: workload = "computers"
c, labels, ne, csr, perm = bench.make_workload("computers")
g = api.VicinityGraph(*csr, device=0)
for s in range(4):
    tg = bench.batch_targets(ne, perm, 500 + s, 0, 1, 1024)
    t0 = time.perf_counter()
    pi, st, cnt = g.vicinity_pi(tg, hop=2, flags=L.F_NORM)
    dt = time.perf_counter() - t0
    ms, nch = g.last_stage_ms()
    print("call %d: wall %.1f ms  chunks %d  stages %s" % (s, dt * 1e3, nch, {k: round(v, 2) for k, v in ms.items()}), flush=True)
print(g.last_counts())
