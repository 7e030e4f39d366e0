#!/usr/bin/env python
"""one production-sized call: every edge of a BASELINE shape (+ equal negatives) through the reference-facing API
   python scripts/full_call.py computers|collab|pubmed [extended]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tlc-gnn_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from tlc_b200 import graphgen as gg
import sg2dgm.riccidist2dgm as mirror
name = sys.argv[1]; ext = len(sys.argv) > 2 and sys.argv[2] == "extended"
hop = int(os.environ.get("HOP", "2"))
c = gg.make_config(name)
labels, ne = gg.relabel_first_appearance(c["edges"])
csr = gg.build_csr(len(labels), ne, c["kappa"])
rng = np.random.default_rng(1)
neg = rng.integers(0, len(labels), size=ne.shape)
tg = np.concatenate([ne, neg]).astype(np.int32)
gm = mirror.graph2pi.from_csr(*csr, device=0)
t0 = time.time()
gm.get_pimg_for_all_edges(tg, cores=16, hop=hop, norm=True, extended_flag=ext, resolution=5, descriptor="sum")
dt = time.time() - t0
cn = gm._graph.last_counts()
print("%s hop %d ext %d: %d targets in %.2f s = %.0f targets/s; computed %d; counts %s" % (name, hop, ext, len(tg), dt, len(tg) / dt, gm.cnt_compute, cn))
# spot check against the oracle
import oracle as orc
og = orc.OracleGraph(*csr)
idx = rng.choice(len(tg), 256, replace=False)
o = og.run_batch(tg[idx], hop=hop, flags=orc.F_NORM | (orc.F_EXTENDED if ext else 0), nthreads=orc.max_threads())
den = np.where(o["pi"] != 0, np.abs(o["pi"]), 1.0)
print("spot check: status equal %s, max rel err %.2e" % (np.array_equal(gm.status[idx], o["status"]), np.max(np.abs(gm.pi_sg[idx] - o["pi"]) / den)))
