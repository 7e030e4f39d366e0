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
