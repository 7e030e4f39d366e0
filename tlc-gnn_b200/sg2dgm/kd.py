"""Mirror of the PDGNN training-set generator's per-node call,
Knowledge_Distillation/data_utils_NC.py:95-183 `compute_persistence_image(g, u, filt='ricci', hop, ricci_curv,
mode='PI')`, and a batched form of it.

Per node the reference returns the 9-tuple
    (dgmOrd0 [n-1,2], dgmExt1 [m-n+1,2], PI [25], filtration_val [n], edge_index [2,m], PI0 [25], PI1 [25],
     PD_time, PI_time)                                                                        (:183)
or `(None, None)` when the ball has no edge (:103-104).  Here all nodes of a call go through ONE C-ABI call
(`tlc_vicinity_detail`, node mode, KD flags: zero-persistence pairs kept, division by max + 1e-10, images of
Ord0 u Ext1 / Ord0 / Ext1).  filt = 'ricci' (the hot path), 'degree', 'centrality' and 'clustering' (:118-128)
and 'hks' (the generators' default; :87-93,115-117, through the Taylor series of the heat kernel instead of an
eigendecomposition, agreement ~1e-13) are computed on the GPU.

Order conventions: the reference's local vertex numbering and pair order follow networkx's sub-graph view iteration
(implementation-defined).  This mirror uses the canonical order: local ids ascending by graph id,
`edge_index` lexicographic (lo, hi), Ord0 / Ext1 in sweep order under that tie-break.  `old_label` gives the graph id
of every local vertex.
"""
import time

import numpy as np

from tlc_b200 import _lib as L

KD_FLAGS = L.F_NORM | L.F_EXTENDED | L.F_KEEP_ZERO | L.F_NORM_EPS


_FILT_FLAGS = {"ricci": 0, "degree": L.F_FILT_DEGREE, "centrality": L.F_FILT_CENTRALITY, "clustering": L.F_FILT_CLUSTERING,
               "hks": L.F_FILT_HKS}   # data_utils_NC.py:115-142 ('hks': set the diffusion time with g2pi._graph.set_hks_time, default 0.1)


def compute_persistence_images(g2pi, nodes, hop=2, resolution=5, as_torch=False, filt="ricci", budget=None):
    """node-centred generator, batched: `g2pi` is a sg2dgm.riccidist2dgm.graph2pi (graph + curvature resident on the
    GPU), `nodes` are ORIGINAL node labels.  Returns a list with one entry per node: the reference's 9-tuple, or
    (None, None)."""
    return _emit(g2pi, [(u, u) for u in nodes], L.MODE_NODE, hop, resolution, as_torch, filt, budget)


def compute_persistence_images_lp(g2pi, pairs, hop=2, resolution=5, as_torch=False, filt="ricci"):
    """edge-centred generator (Knowledge_Distillation/data_utils_LP.py:105-196), batched over (u, v) pairs of ORIGINAL
    labels: vicinity = (ball(u) & ball(v)) + [u] + [v] (:111), filtration = distance to both roots (:35-60).  A disconnected
    vicinity is outside the contract (the reference has no check and fails or not by accident): (None, None) here."""
    return _emit(g2pi, [(u, v) for u, v in pairs], L.MODE_EDGE_FORCED, hop, resolution, as_torch, filt)


EMIT_BUDGET = 8_000_000   # vertices + edges of the vicinities handed to ONE tlc_vicinity_detail call (every intermediate of
                          # a call comes back to the host: ~100 bytes per simplex)


def _batches(G, tg, hop, mode, budget):
    """split a target list into consecutive batches whose vicinities hold at most `budget` simplices in total"""
    n, m, _ = G.vicinity_sizes(tg, hop=hop, mode=mode)
    cost = n.astype(np.int64) + m.astype(np.int64) + 1
    out, start, acc = [], 0, 0
    for i, c in enumerate(cost.tolist()):
        if i > start and acc + c > budget:
            out.append((start, i))
            start, acc = i, 0
        acc += c
    if start < len(tg):
        out.append((start, len(tg)))
    return out


def _emit(g2pi, targets, mode, hop, resolution, as_torch, filt="ricci", budget=None):
    if filt not in _FILT_FLAGS:
        raise NotImplementedError("filt=%r: only 'ricci', 'degree', 'centrality', 'clustering' are on the GPU path (SURVEY.md row N3)" % (filt,))
    tg = g2pi._map_targets(targets)
    G = g2pi._graph
    out = []
    for lo, hi in _batches(G, tg, hop, mode, EMIT_BUDGET if budget is None else budget):
        out += _emit_batch(g2pi, G, tg[lo:hi], mode, hop, resolution, as_torch, filt)
    return out


def _emit_batch(g2pi, G, tg, mode, hop, resolution, as_torch, filt):
    targets = tg
    t0 = time.time()
    d = G.vicinity_detail(tg, hop=hop, mode=mode, descriptor="sum", resolution=resolution, flags=KD_FLAGS | _FILT_FLAGS[filt])
    dt = (time.time() - t0) / max(1, len(targets))
    out = []
    for i in range(len(targets)):
        a = G.per_target(d, i)
        if a["status"] != L.ST_OK:           # lone centre / unknown node: `return None, None`   :103-104
            out.append((None, None))
            continue
        pk = a["pkind"]
        pairs = np.stack([a["pbirth"], a["pdeath"]], 1)
        ord0, ext1 = pairs[pk == L.K_UP], pairs[pk == L.K_ONE]
        edge_index = np.stack([a["elo"], a["ehi"]], 0).astype(np.int64)
        if as_torch:
            import torch
            edge_index = torch.from_numpy(edge_index)                                           # :109
        old = a["vert"] if g2pi.old_label is None else [g2pi.old_label[int(x)] for x in a["vert"]]
        tup = (ord0, ext1, a["img"].copy(), list(a["fval"]), edge_index, a["img_up"].copy(), a["img_one"].copy(), dt, 0.0)
        out.append(_KDTuple(tup, old))
    return out


class _KDTuple(tuple):
    """the reference's 9-tuple, plus `.old_label` (graph label of every local vertex)."""

    def __new__(cls, items, old_label):
        self = super().__new__(cls, items)
        self.old_label = old_label
        return self


def compute_persistence_image(g2pi, u, v=None, filt="ricci", hks_time=0.1, hop=2, ricci_curv=None, mode="PI", **_unused):
    """per-target signatures of data_utils_NC.py:95 (g, u, ...) and data_utils_LP.py:105 (g, u, v, ...); the first
    argument is the graph2pi object that holds graph and curvature."""
    if mode != "PI":
        raise NotImplementedError("only mode='PI' is on the GPU path")
    if filt == "hks":
        g2pi._graph.set_hks_time(hks_time)
    if v is None:
        return compute_persistence_images(g2pi, [u], hop=hop, filt=filt)[0]
    return compute_persistence_images_lp(g2pi, [(u, v)], hop=hop, filt=filt)[0]


def compute_persistence_images_gc(graphs, filt="degree", resolution=5, device=0):
    """graph-classification generator (Knowledge_Distillation/data_utils_GC.py:95-167), batched: every WHOLE graph is one
    vicinity.  graphs: list of (n, edges[m, 2]) with nodes 0..n-1 (the TU-dataset numbering the reference's edge_index
    assumes).  All graphs go into one block-diagonal CSR and ONE C-ABI call (node mode from node 0 with hop >= the largest
    graph, i.e. the ball is node 0's component); a graph without edges or not connected yields (None, None) as
    `if len(subgraph.edges()) == 0 or not nx.is_connected(subgraph)` does (:99-100).
    filt: 'degree' / 'centrality'.  (filt='ricci' is not mirrored: as written, :118-122 indexes the curvature LIST with
    a tuple inside a try/except, so every distance becomes 100 -- or it raises when handed a dict.)"""
    import numpy as np
    from tlc_b200 import api
    from tlc_b200.graphgen import build_csr
    if filt not in ("degree", "centrality"):
        raise NotImplementedError("data_utils_GC mirror: filt must be 'degree' or 'centrality'")
    sizes = [int(n) for n, _ in graphs]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    es = [np.asarray(e, dtype=np.int64).reshape(-1, 2) + offs[k] for k, (_, e) in enumerate(graphs)]
    E = np.concatenate(es) if es else np.zeros((0, 2), np.int64)
    E = E[E[:, 0] != E[:, 1]]
    lo, hi = np.minimum(E[:, 0], E[:, 1]), np.maximum(E[:, 0], E[:, 1])
    key = np.unique(lo * int(offs[-1] + 1) + hi)
    E = np.stack([key // int(offs[-1] + 1), key % int(offs[-1] + 1)], 1)
    G = api.VicinityGraph(*build_csr(int(offs[-1]), E, np.zeros(len(E))), device=device)
    tg = np.stack([offs[:-1], offs[:-1]], 1).astype(np.int32)
    t0 = time.time()
    d = G.vicinity_detail(tg, hop=max(sizes + [1]), mode=L.MODE_NODE, descriptor="sum", resolution=resolution,
                          flags=KD_FLAGS | _FILT_FLAGS[filt])
    dt = (time.time() - t0) / max(1, len(graphs))
    out = []
    for k in range(len(graphs)):
        a = G.per_target(d, k)
        if a["status"] != L.ST_OK or a["n"] != sizes[k]:      # no edge / isolated node 0 / more than one component
            out.append((None, None))
            continue
        pk = a["pkind"]
        pairs = np.stack([a["pbirth"], a["pdeath"]], 1)
        edge_index = np.stack([a["elo"], a["ehi"]], 0).astype(np.int64)
        out.append(_KDTuple((pairs[pk == L.K_UP], pairs[pk == L.K_ONE], a["img"].copy(), list(a["fval"]), edge_index,
                             a["img_up"].copy(), a["img_one"].copy(), dt, 0.0), a["vert"] - offs[k]))
    G.close()
    return out
