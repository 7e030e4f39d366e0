"""shared test helpers: fixtures -> CSR in the reference's numbering, oracle drivers."""
import os

import numpy as np

from tlc_b200 import graphgen as gg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAPH_CASES = ["toy7_hop2", "toy7_hop1", "cora_q_hop2", "pubmed_s_hop2_dyadic", "pubmed_s_hop2_cont",
               "computers_s_hop1", "computers_s_hop2", "pubmed_s_min", "pubmed_s_max"]


def load_case(tag):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    c = {k: z[k] for k in z.files}
    c["hop"] = int(c["hop"])
    c["descriptor"] = str(c["descriptor"])
    labels, ne = gg.relabel_first_appearance(c["edges"])
    c["labels"], c["new_edges"] = labels, ne
    c["csr"] = gg.build_csr(len(labels), ne, c["kappa"])
    lut = {int(l): i for i, l in enumerate(labels)}
    c["new_targets"] = np.array([[lut.get(int(a), -1), lut.get(int(b), -1)] for a, b in c["targets"]], dtype=np.int32)
    return c


def seg(c, name, i):
    off = c["stage_%s_off" % name]
    return c["stage_%s" % name][off[i]:off[i + 1]]


def rel_err(a, ref):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    den = np.where(ref != 0, np.abs(ref), 1.0)
    return float(np.max(np.abs(a - ref) / den)) if a.size else 0.0


KD_CASES = ["kd_ppi_s_hop1", "kd_pubmed_s_hop2_cont", "kd_pubmed_s_hop0", "kd_ppi_s_hop1_degree", "kd_ppi_s_hop1_centrality",
            "kd_ppi_s_hop1_clustering"]
KD_LP_CASES = ["kd_lp_pubmed_s_hop2_cont", "kd_lp_ppi_s_hop1"]  # edge-centred generator (data_utils_LP.py), nodes = [pairs, 2]


def load_kd_case(tag):
    """PDGNN generator fixture (oracle/make_golden.py: make_kd_case): per node the unmodified reference's 9-tuple."""
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    c = {k: z[k] for k in z.files}
    c["hop"] = int(c["hop"])
    c["filt_name"] = str(c["filt_name"]) if "filt_name" in c else "ricci"
    labels, ne = gg.relabel_first_appearance(c["edges"])
    c["csr"] = gg.build_csr(len(labels), ne, c["kappa"])
    c["lut"] = {int(l): i for i, l in enumerate(labels)}
    c["new_nodes"] = np.array([c["lut"][int(u)] for u in c["nodes"].reshape(-1)], dtype=np.int32).reshape(c["nodes"].shape)
    return c


def kd_seg(c, name, k):
    off = c["kd_%s_off" % name]
    w = 2 if name in ("ord0", "ext1", "edge_index") else 1  # offsets count rows; these are [rows, 2] flattened
    return c["kd_%s" % name][w * off[k]:w * off[k + 1]]


def sorted_rows(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 2)
    return a[np.lexsort((a[:, 1], a[:, 0]))]


def kd_expected(c, k):
    """node k of a KD fixture in canonical form: graph ids (first-appearance numbering) ascending, the filtration
    values in that order, the induced edge set as sorted (lo, hi) graph-id pairs, Ord0 / Ext1 as sorted multisets
    (the reference's own vertex / pair order is implementation-defined)."""
    old = kd_seg(c, "old_label", k)
    newid = np.array([c["lut"][int(x)] for x in old], dtype=np.int64)
    order = np.argsort(newid)
    ei = kd_seg(c, "edge_index", k).reshape(-1, 2)
    eg = newid[ei] if len(ei) else np.zeros((0, 2), np.int64)
    eg = np.stack([eg.min(1), eg.max(1)], 1) if len(eg) else eg
    eg = eg[np.lexsort((eg[:, 1], eg[:, 0]))] if len(eg) else eg
    return dict(vert=newid[order], filt=kd_seg(c, "filt", k)[order], edges=eg,
                ord0=sorted_rows(kd_seg(c, "ord0", k)), ext1=sorted_rows(kd_seg(c, "ext1", k)),
                pi=c["pi"][k], pi0=c["pi0"][k], pi1=c["pi1"][k], none=bool(c["none"][k]))


def load_kd_gc_case(tag="kd_gc_degree"):
    """graph-classification generator fixture: list of (n, edges) + the unmodified reference's per-graph outputs."""
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    c = {k: z[k] for k in z.files}
    graphs = []
    for k, n in enumerate(c["sizes"]):
        off = c["kd_gedges_off"]
        graphs.append((int(n), c["kd_gedges"][2 * off[k]:2 * off[k + 1]].reshape(-1, 2)))
    c["graphs"] = graphs
    c["filt_name"] = str(c["filt_name"])
    return c
