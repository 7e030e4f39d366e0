"""the vicinity shapes of the legacy `sg2pimg(range=...)` (riccidist2dgm.py:242-247 'union', :289-296 'removeinter'): the
oracle's vertex sets against the reference's own lines executed with networkx (CPU)."""
import networkx as nx
import numpy as np

import oracle as orc
from tlc_b200 import graphgen as gg


def test_union_and_removeinter_vertex_sets_match_the_reference_lines():
    c = gg.make_config("pubmed", scale=0.1)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    csr = gg.build_csr(len(labels), ne, c["kappa"])
    G = nx.Graph()
    G.add_nodes_from(range(len(labels)))
    G.add_edges_from(ne.tolist())
    og = orc.OracleGraph(*csr)
    rng = np.random.default_rng(0)
    tg = np.concatenate([ne[rng.choice(len(ne), 30, replace=False)], rng.integers(0, len(labels), size=(10, 2))])
    for u, v in tg:
        u, v = int(u), int(v)
        if u == v or G.degree(u) == 0 or G.degree(v) == 0:
            continue
        for hop in (1, 2):
            nodes_u = [u] + [x for _, x in nx.bfs_edges(G, u, depth_limit=hop)]          # :243
            nodes_v = [v] + [x for _, x in nx.bfs_edges(G, v, depth_limit=hop)]          # :245
            union = sorted(set(nodes_u + nodes_v))                                       # :246 (G.subgraph dedups)
            inter = set(nodes_u) & set(nodes_v)                                          # :294
            rem = sorted(set(list(set(nodes_u + nodes_v).difference(inter)) + [u, v]))   # :295
            for mode, ref in ((orc.MODE_EDGE_UNION, union), (orc.MODE_EDGE_REMOVEINTER, rem)):
                d = og.run_one(u, v, hop=hop, mode=mode, flags=orc.F_NORM)
                assert d["vert"].tolist() == ref, (u, v, hop, mode)
                assert d["lu"] >= 0 and d["lv"] >= 0    # both roots are members of either shape
