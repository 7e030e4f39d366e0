"""BASELINE.json's configurations at FULL size: the CUDA path (through the C-ABI) checked by size-independent
properties of the domain plus oracle spot checks on a sample the CPU finishes in seconds:
  * symmetry: the vicinity, the 'sum' filtration and hence diagram and image of (u, v) and (v, u) are identical;
  * order independence: a permuted target list gives the permuted rows (chunk planning, scatter);
  * route independence: graph-row route == materialised-adjacency route, bit for bit;
  * the oracle on a seeded sample: status, counts, images (1e-5), and -- through the stage-level entry point --
    filtration values and persistence pairs bit-exact.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as orc
from helpers import rel_err
from tlc_b200 import _lib as L
from tlc_b200 import api, graphgen as gg

IMG_TOL = 1e-5


def full_graph(name, **kw):
    c = gg.make_config(name, **kw)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    csr = gg.build_csr(len(labels), ne, c["kappa"])
    return c, labels, ne, csr


def check_pairs_bit_exact(g, og, tg, hop, flags, oflags, mode=L.MODE_EDGE):
    d = g.vicinity_detail(tg, hop=hop, mode=mode, flags=flags)
    for i, (u, v) in enumerate(tg):
        a = g.per_target(d, i)
        o = og.run_one(int(u), int(v), hop=hop, mode=mode, flags=oflags)
        assert a["status"] == o["status"] and a["n"] == o["n"]
        if o["status"] > 1:
            continue
        assert np.array_equal(a["vert"], o["vert"]) and np.array_equal(a["fval"], o["fval"])
        keep = np.ones(len(o["pkind"]), bool) if not (flags & L.F_ASC_ONLY) else np.isin(o["pkind"], (L.K_UP, L.K_ESS))
        assert np.array_equal(a["pkind"], o["pkind"][keep])
        assert np.array_equal(a["pbv"], o["pbv"][keep]) and np.array_equal(a["pdv"], o["pdv"][keep])
        assert np.array_equal(a["pbirth"], o["pbirth"][keep]) and np.array_equal(a["pdeath"], o["pdeath"][keep])


def test_computers_full_2hop():
    """configs[2]: Computers-shaped (13,752 nodes / 245,861 edges), dense 2-hop vicinities."""
    c, labels, ne, csr = full_graph("computers")
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    rng = np.random.default_rng(42)
    tg = ne[rng.choice(len(ne), 384, replace=False)].astype(np.int32)
    both = np.concatenate([tg, tg[:, ::-1]])
    pi, st, cnt = g.vicinity_pi(both, hop=2, flags=L.F_NORM)
    assert g.last_counts()["graph_row_route"] > 0 and g.last_counts()["handed_back"] == 0
    assert np.array_equal(st[:384], st[384:]) and np.array_equal(pi[:384], pi[384:])       # symmetry, bit-exact
    perm = rng.permutation(len(both))
    pi_p, st_p, cnt_p = g.vicinity_pi(both[perm], hop=2, flags=L.F_NORM)
    assert cnt_p == cnt and np.array_equal(st_p, st[perm]) and np.array_equal(pi_p, pi[perm])  # order independence
    pi_m, st_m, cnt_m = g.vicinity_pi(both, hop=2, flags=L.F_NORM | L.F_NO_DIRECT)
    assert g.last_counts()["graph_row_route"] == 0
    assert cnt_m == cnt and np.array_equal(st_m, st) and np.array_equal(pi_m, pi)          # route independence
    o = og.run_batch(tg[:48], hop=2, flags=orc.F_NORM, nthreads=orc.max_threads())
    assert np.array_equal(st[:48], o["status"]) and rel_err(pi[:48], o["pi"]) < IMG_TOL
    check_pairs_bit_exact(g, og, tg[:3], 2, L.F_NORM | L.F_DIRECT | L.F_ASC_ONLY, orc.F_NORM)
    check_pairs_bit_exact(g, og, tg[3:5], 2, L.F_NORM | L.F_EXTENDED, orc.F_NORM | orc.F_EXTENDED)
    # hop 1 = the reference's own setting for this dataset (baselines/TLCGNN.py:102), extended_flag=True (loaddatas.py:100)
    t1 = ne[rng.choice(len(ne), 4096, replace=False)].astype(np.int32)
    pi1, st1, cnt1 = g.vicinity_pi(t1, hop=1, flags=L.F_NORM | L.F_EXTENDED)
    o1 = og.run_batch(t1, hop=1, flags=orc.F_NORM | orc.F_EXTENDED, nthreads=orc.max_threads())
    assert cnt1 == o1["cnt_compute"] and np.array_equal(st1, o1["status"]) and rel_err(pi1, o1["pi"]) < IMG_TOL
    g.close()


def test_pubmed_full_2hop_extended():
    """configs[1]: PubMed-shaped (19,717 / 44,324), 2-hop, Ricci filtration; every edge in ONE call, extended_flag=True
    (the reference's shipped setting for PubMed), against the oracle on all of them."""
    c, labels, ne, csr = full_graph("pubmed")
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    tg = ne.astype(np.int32)
    for ext in (0, 1):
        fl = L.F_NORM | (L.F_EXTENDED if ext else 0)
        pi, st, cnt = g.vicinity_pi(tg, hop=2, flags=fl)
        o = og.run_batch(tg, hop=2, flags=orc.F_NORM | (orc.F_EXTENDED if ext else 0), nthreads=orc.max_threads())
        assert cnt == o["cnt_compute"] and np.array_equal(st, o["status"])
        assert rel_err(pi, o["pi"]) < IMG_TOL
    g.close()


def test_collab_full_2hop_with_negatives():
    """configs[4]: collab-shaped (235,868 / 1,285,465), edges + equal negatives, 2-hop; a 16,384-target sample."""
    c, labels, ne, csr = full_graph("collab")
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    rng = np.random.default_rng(5)
    pos = ne[rng.choice(len(ne), 8192, replace=False)]
    neg = rng.integers(0, len(labels), size=(8192, 2))
    tg = np.concatenate([pos, neg]).astype(np.int32)
    pi, st, cnt = g.vicinity_pi(tg, hop=2, flags=L.F_NORM)
    o = og.run_batch(tg, hop=2, flags=orc.F_NORM, nthreads=orc.max_threads())
    assert cnt == o["cnt_compute"] and np.array_equal(st, o["status"])
    assert rel_err(pi, o["pi"]) < IMG_TOL
    pi_r, st_r, _ = g.vicinity_pi(tg[:, ::-1].copy(), hop=2, flags=L.F_NORM)
    assert np.array_equal(st_r, st) and np.array_equal(pi_r, pi)                            # symmetry
    n, m, _ = g.vicinity_sizes(tg)
    big = np.argsort(-m)[:4]                                                                # the heaviest vicinities of the sample
    check_pairs_bit_exact(g, og, tg[big], 2, L.F_NORM | L.F_EXTENDED, orc.F_NORM | orc.F_EXTENDED)
    g.close()


def test_ppi_24_graphs_node_mode():
    """configs[3]: 24 PPI-shaped graphs (block-diagonal CSR), PDGNN node-centred 2-hop vicinities, KD flags."""
    c, labels, ne, csr = full_graph("ppi", n_graphs=24)
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    rng = np.random.default_rng(9)
    nodes = rng.choice(len(labels), 96, replace=False)
    tg = np.stack([nodes, nodes], 1).astype(np.int32)
    fl = L.F_NORM | L.F_EXTENDED | L.F_KEEP_ZERO | L.F_NORM_EPS
    pi, st, cnt = g.vicinity_pi(tg, hop=2, mode=L.MODE_NODE, flags=fl)
    o = og.run_batch(tg, hop=2, mode=orc.MODE_NODE, flags=fl, nthreads=orc.max_threads())
    assert cnt == o["cnt_compute"] and np.array_equal(st, o["status"]) and rel_err(pi, o["pi"]) < IMG_TOL
    # a vicinity never leaves its graph: every vertex of node u's ball lies in u's block of 2,400 ids
    d = g.vicinity_detail(tg[:4], hop=2, mode=L.MODE_NODE, flags=fl)
    per = len(labels) // 24
    for i in range(4):
        a = g.per_target(d, i)
        lab = labels[a["vert"]]
        assert (lab // (c["N"] // 24) == labels[nodes[i]] // (c["N"] // 24)).all()
    check_pairs_bit_exact(g, og, tg[:3], 2, fl, fl, mode=L.MODE_NODE)
    # ascending-only variant through the graph-row route
    fl0 = L.F_NORM | L.F_KEEP_ZERO | L.F_NORM_EPS
    pi_d, st_d, _ = g.vicinity_pi(tg, hop=2, mode=L.MODE_NODE, flags=fl0 | L.F_DIRECT)
    pi_m, st_m, _ = g.vicinity_pi(tg, hop=2, mode=L.MODE_NODE, flags=fl0 | L.F_NO_DIRECT)
    assert np.array_equal(pi_d, pi_m) and np.array_equal(st_d, st_m)
    g.close()
