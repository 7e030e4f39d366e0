"""the GPU-resident feature table + .npy cache against what the reference's caller code does on
the host (loaddatas.py:56-103, baselines/TLCGNN.py:35-53) -- bit-exact (float64 -> float32 rounding is torch's)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from helpers import load_case


def host_decode_rows(PI, s, kind, index=None):
    """baselines/TLCGNN.py:35-53 verbatim semantics on the host table."""
    import torch
    tp, tn, vp, vn = s
    if kind == "train":
        PIs = np.concatenate((PI[:tp], PI[tp:tp + tn][index]))
    elif kind == "val":
        PIs = PI[tp + tn:tp + tn + vp + vn]
    else:
        PIs = PI[tp + tn + vp + vn:]
    return torch.Tensor(PIs.reshape((len(PIs), -1)))


def test_table_rows_match_host_decode(tmp_path):
    import torch
    from tlc_b200.table import PITable, cache_filename
    rng = np.random.default_rng(0)
    E, r2 = 5000, 25
    PI = rng.random((E, r2)) * rng.choice([1e-12, 1.0, 37.5], size=(E, 1))
    splits = [2000, 2300, 150, 150, 200, 200]
    t = PITable(PI, splits=splits)
    index = np.random.randint(0, splits[1], splits[0])                    # TLCGNN.py:38
    for kind in ("train", "val", "test"):
        got = t.rows(kind, index=index if kind == "train" else None)
        assert got.is_cuda and got.dtype == torch.float32
        ref = host_decode_rows(PI, splits[:4], kind, index)
        assert torch.equal(got.cpu(), ref)
    with pytest.raises(IndexError):
        t.rows("train", index=np.array([splits[1]]))
    from tlc_b200 import _lib as L
    with pytest.raises(L.TlcError):
        t.gather(index=np.array([E]))
    # .npy cache round trip in the reference's layout
    fn = cache_filename("computers", str(tmp_path))
    assert fn.endswith("Computers.npy")
    t.save_npy(fn)
    back = np.load(fn)
    assert back.dtype == np.float64 and np.array_equal(back, PI)
    t2 = PITable.from_npy(fn, splits=splits)
    assert torch.equal(t2.rows("test").cpu(), host_decode_rows(PI, splits[:4], "test"))


def test_compute_persistence_image_cache(tmp_path):
    """loaddatas.compute_persistence_image mirror: miss -> compute + np.save, hit -> np.load; rows == reference fixture."""
    import networkx as nx
    import sg2dgm.feature_cache as fc
    c = load_case("cora_q_hop2")
    g = nx.Graph()
    g.add_edges_from([(int(a), int(b)) for a, b in c["edges"]])
    ricci = sorted([[int(a), int(b), float(k)] for (a, b), k in zip(c["edges"], c["kappa"])] +
                   [[int(b), int(a), float(k)] for (a, b), k in zip(c["edges"], c["kappa"])])
    tg = c["targets"]
    parts = [tg[:60], tg[60:100], tg[100:110], tg[110:120], tg[120:135], tg[135:]]
    pi_sg, table = fc.compute_persistence_image(g, ricci, *parts, data_name="toy", hop=c["hop"], cache_dir=str(tmp_path))
    ref = c["pi_ext1"]                                                     # the REAL reference's pi_sg (extended_flag=True)
    den = np.where(ref != 0, np.abs(ref), 1.0)
    assert np.max(np.abs(pi_sg - ref) / den) < 1e-5
    assert os.path.exists(os.path.join(str(tmp_path), "toy.npy"))
    pi2, table2 = fc.compute_persistence_image(g, ricci, *parts, data_name="toy", hop=c["hop"], cache_dir=str(tmp_path))
    assert np.array_equal(pi2, pi_sg)
    import torch
    assert torch.equal(table.rows("val"), table2.rows("val"))
    assert torch.equal(table.rows("val").cpu(), torch.Tensor(pi_sg[100:120]))
