#!/usr/bin/env python
"""Launch EVERY kernel of libtlc_b200.so at a representative size, a few times each, so that one
`ncu --set full` pass (or a launch list) covers the whole library:

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_all python scripts/ncu_all_kernels.py

Workloads: Computers-shaped 2-hop (graph-row route: ball cache, light sizes, filtration<1,1>, vertex order, sweep, image),
the same with extended_flag (materialised route: sizes, fill, filtration<0,1>, edge list, sort, union-find, loops, image),
PubMed-shaped extended with and without kernel S (lane-per-vicinity loops kernels / the three classes of kernel S),
the PDGNN structural filtrations (degree, clustering), the decoder hand-off (gather), and the peer-store exchange
(a one-rank table: the same scatter + wait kernels).  Inputs are fresh per call; nothing here is timed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tlc-gnn_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from tlc_b200 import _lib as L  # noqa: E402
from tlc_b200 import api  # noqa: E402
from tlc_b200.table import PITable  # noqa: E402


def main():
    small = "small" in sys.argv[1:]                         # smaller batches (sanitizer runs)
    only = [a[5:].split(",") for a in sys.argv[1:] if a.startswith("only=")]   # only=computers,pubmed,cora,ricci,gather
    want = lambda sec: not only or sec in only[0]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    # ---- Computers-shaped ----
    if want("computers"):
        section_computers(small, dev)
    if want("pubmed"):
        section_pubmed(small)
    if want("cora"):
        section_cora(small)
    if want("ricci"):
        section_ricci(small)
    if want("gather"):
        section_gather(dev)
    print("all kernels launched: %d launches" % api.launch_count())


def section_computers(small, dev):
    c, labels, ne, csr, perm = bench.make_workload("computers")
    g = api.VicinityGraph(*csr, device=0)
    B = 64 if small else 512
    for s in range(2):
        tg = bench.batch_targets(ne, perm, s, 0, 1, B)
        g.vicinity_pi(tg, hop=2, flags=L.F_NORM)                               # graph-row route
    tg = bench.batch_targets(ne, perm, 3, 0, 1, B)
    g.vicinity_pi(tg, hop=2, flags=L.F_NORM | L.F_NO_TABLE)                    # phased Dijkstra (kernel 1b) instead of the SSSP tables
    tg = bench.batch_targets(ne, perm, 7, 0, 1, 16 if small else 32)
    g.vicinity_pi(tg, hop=2, flags=L.F_NORM | L.F_EXTENDED)                    # materialised route + loops (CTA per vicinity)
    g.vicinity_pi(tg, hop=2, flags=L.F_NORM | L.F_EDGE_SORTED | L.F_NO_DIRECT)  # edge-sorted ascending sweep
    tg = bench.batch_targets(ne, perm, 9, 0, 1, 512 if small else 4096)
    g.vicinity_pi(tg, hop=1, flags=L.F_NORM | L.F_EXTENDED)                    # kernel S on dense 1-hop vicinities
    # peer-store exchange with a single rank (same kernels as N > 1)
    ptr, handle = g.table_create(2 * 4096, 5)
    g.table_attach([handle], 0)
    t_dev = torch.from_numpy(tg).to(dev)
    rows = torch.arange(len(tg), dtype=torch.int64, device=dev)
    g.vicinity_pi_exchange(t_dev, rows, hop=1, flags=L.F_NORM)
    torch.cuda.synchronize()
    g.close()


def section_pubmed(small):
    c, labels, ne, csr, perm = bench.make_workload("pubmed")
    g = api.VicinityGraph(*csr, device=0)
    B = 1024 if small else 8192
    tg = bench.batch_targets(ne, perm, 0, 0, 1, B)
    g.vicinity_pi(tg, hop=2, flags=L.F_NORM | L.F_EXTENDED)                    # kernel S, three classes
    g.vicinity_pi(tg, hop=2, flags=L.F_NORM | L.F_EXTENDED | L.F_NO_SMALL)      # staged: lane-per-vicinity loops kernels
    g.vicinity_pi(tg, hop=2, flags=L.F_NORM | L.F_NO_SMALL)
    ids = np.arange(256 if small else 2048, dtype=np.int32)
    nodes = np.stack([ids, ids], 1)
    kd = L.F_NORM | L.F_EXTENDED | L.F_KEEP_ZERO | L.F_NORM_EPS
    g.vicinity_pi(nodes, hop=2, mode=L.MODE_NODE, flags=kd | L.F_FILT_DEGREE)       # PDGNN structural filtrations
    g.vicinity_pi(nodes, hop=2, mode=L.MODE_NODE, flags=kd | L.F_FILT_CLUSTERING)
    g.set_hks_time(0.1)
    g.vicinity_pi(nodes, hop=2, mode=L.MODE_NODE, flags=kd | L.F_FILT_HKS)            # heat kernel signature
    tg = bench.batch_targets(ne, perm, 5, 0, 1, 256 if small else 2048)
    g.vicinity_pi(tg, hop=1, mode=L.MODE_EDGE_UNION, flags=L.F_NORM | L.F_EXTENDED)   # legacy vicinity shapes
    g.vicinity_pi(tg, hop=1, mode=L.MODE_EDGE_REMOVEINTER, flags=L.F_NORM | L.F_EXTENDED)
    g.close()


def section_ricci(small):
    """Ollivier-Ricci curvature of every edge (kernel 6)"""
    n_r = 600 if small else 3000
    rng = np.random.default_rng(5)
    from tlc_b200.graphgen import build_csr
    e = rng.integers(0, n_r, size=(4 * n_r, 2))
    e = e[e[:, 0] != e[:, 1]]
    key = np.unique(np.minimum(e[:, 0], e[:, 1]) * n_r + np.maximum(e[:, 0], e[:, 1]))
    rp, cl, _ = build_csr(n_r, np.stack([key // n_r, key % n_r], 1), np.zeros(len(key)))
    api.ollivier_ricci(rp, cl, device=0)


def section_cora(small):
    """Cora-shaped, hop-distance filtration: hundreds of exact ties -> kernel 3v hands targets back (redo + scatter)"""
    c, labels, ne, csr, perm = bench.make_workload("cora")
    g = api.VicinityGraph(*csr, device=0)
    tg = bench.batch_targets(ne, perm, 0, 0, 1, 256 if small else 2048)
    g.vicinity_pi(tg, hop=2, flags=L.F_NORM | L.F_DIRECT)
    g.vicinity_pi(tg, hop=3, flags=L.F_NORM)
    g.close()


def section_gather(dev):
    """decoder hand-off"""
    E = 100000
    table = PITable(torch.rand((E, 25), dtype=torch.float64, device=dev), splits=[60000, 30000, 2500, 2500, 2500, 2500])
    idx = torch.randint(0, E, (50000,), device=dev)
    out = torch.empty((50000, 25), dtype=torch.float32, device=dev)
    table.gather(index=idx, out=out)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
