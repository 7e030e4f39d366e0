"""Mirror of the caller of the hot path, loaddatas.py:56-103 `compute_persistence_image`, without the dataset
handling around it (PyG loading, edge split and Ricci curvature are inputs here):

    cache hit  -> np.load('./data/TLCGNN/<Name>.npy')                                  (:62-64)
    cache miss -> graph2pi(g, ricci_curv).get_pimg_for_all_edges(total_edges, cores=16, hop=hop, norm=True,
                  extended_flag=True, resolution=5, descriptor='sum'); np.save(filename, pi.pi_sg)   (:99-102)

and returns the table both as the reference's host array and as a GPU-resident PITable for the decoder.
"""
import os

import numpy as np

from tlc_b200.table import PITable, cache_filename

from . import riccidist2dgm as sg2dgm


def compute_persistence_image(g, ricci_cur, train_edges, train_edges_false, val_edges, val_edges_false, test_edges,
                              test_edges_false, data_name, hop=1, cache_dir="./data/TLCGNN", device=0,
                              extended_flag=True):
    """returns (pi_sg float64[E, 25], PITable).  `g`: the training graph (val/test positives already removed,
    loaddatas.py:71-92), `ricci_cur`: [[n1, n2, kappa], ...] (:117-121)."""
    parts = [np.asarray(p).reshape(-1, 2) for p in
             (train_edges, train_edges_false, val_edges, val_edges_false, test_edges, test_edges_false)]
    splits = [len(p) for p in parts]
    filename = cache_filename(data_name, cache_dir)
    if os.path.exists(filename):                                  # :63-64
        pi_sg = np.load(filename)
        return pi_sg, PITable(pi_sg, splits=splits if sum(splits) == len(pi_sg) else None, device=device)
    total_edges = np.concatenate(parts)                           # :65-66
    pi = sg2dgm.graph2pi(g, ricci_curv=ricci_cur, device=device)
    pi.get_pimg_for_all_edges(total_edges, cores=16, hop=hop, norm=True, extended_flag=extended_flag,
                              resolution=5, descriptor='sum')     # :99-101
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    np.save(filename, pi.pi_sg)                                   # :102
    return pi.pi_sg, PITable(pi.pi_sg, splits=splits, device=device)


def compute_ricci_curvature(data=None, edge_index=None, alpha=0.5, device=0):
    """Mirror of loaddatas.py:105-123: Ollivier-Ricci curvature (alpha-lazy measures, Sinkhorn transport) of every edge of
    the graph given by `data.edge_index` (or `edge_index`, a [2, E] array of node labels) -- computed by kernel 6
    (tlc_ollivier_ricci) instead of GraphRicciCurvature + POT.  Returns the reference's `ricci_list`: [[n1, n2, kappa],
    [n2, n1, kappa], ...] sorted, ready for graph2pi(g, ricci_curv=...).  As nx.Graph does, duplicate edges collapse and
    self-loops are dropped (OllivierRicci removes them)."""
    from tlc_b200 import api
    from tlc_b200.graphgen import build_csr
    ei = np.asarray(edge_index if edge_index is not None else data.edge_index)
    ei = ei.reshape(2, -1)
    labels, inv = np.unique(ei.reshape(-1), return_inverse=True)       # sorted labels -> dense ids
    e = inv.reshape(2, -1).T.astype(np.int64)
    e = e[e[:, 0] != e[:, 1]]
    N = len(labels)
    key = np.unique(np.minimum(e[:, 0], e[:, 1]) * N + np.maximum(e[:, 0], e[:, 1]))
    und = np.stack([key // N, key % N], 1)
    rowptr, col, _ = build_csr(N, und, np.zeros(len(und)))
    kap = api.ollivier_ricci(rowptr, col, alpha=alpha, device=device)
    out = []
    for x in range(N):
        for q in range(rowptr[x], rowptr[x + 1]):
            out.append([labels[x].item(), labels[col[q]].item(), float(kap[q])])
    return sorted(out)
