"""The CPU restatement against the REAL reference imported in place from /root/reference (build container
only; the tree does not exist on the GPU box -> skipped there).  Fresh random graphs each run seed, i.e.
inputs that are NOT in the committed fixtures: as-is batch output (images, counts) and the canonical-order
stage outputs (filtration values, PD0, Pos/Neg, PD1) must agree bit for bit / to 1e-11."""
import numpy as np
import pytest

import oracle as orc
import ref_harness as rh
from helpers import rel_err
from tlc_b200 import graphgen as gg

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not rh.available(), reason="reference tree not mounted")]


def _case(name, scale, cont, seed, ntargets):
    c = gg.make_config(name, scale=scale, continuous=cont)
    rng = np.random.default_rng(seed)
    e = c["edges"]
    tg = e[rng.choice(len(e), ntargets, replace=False)]
    far = rng.integers(0, c["N"], size=(4, 2))
    return c, np.concatenate([tg, far])


@pytest.mark.parametrize("name,scale,hop,cont,ext", [("cora", 0.25, 2, False, True), ("pubmed", 0.04, 2, True, True),
                                                     ("pubmed", 0.04, 2, False, False), ("computers", 0.02, 1, True, True)])
def test_batch_matches_reference(name, scale, hop, cont, ext):
    c, tg = _case(name, scale, cont, 11, 24)
    pi_ref, cnt_ref, _ = rh.run_batch(c["edges"], c["kappa"], tg, hop, ext)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    lut = {int(l): i for i, l in enumerate(labels)}
    new_t = np.array([[lut.get(int(a), -1), lut.get(int(b), -1)] for a, b in tg], dtype=np.int32)
    og = orc.OracleGraph(*gg.build_csr(len(labels), ne, c["kappa"]))
    r = og.run_batch(new_t, hop=hop, flags=orc.F_NORM | (orc.F_EXTENDED if ext else 0))
    assert r["cnt_compute"] == cnt_ref
    assert np.array_equal(pi_ref.any(axis=1), r["pi"].any(axis=1))
    assert rel_err(r["pi"], pi_ref) < 1e-11


def test_stages_match_canonical_reference():
    c, tg = _case("pubmed", 0.04, True, 5, 10)
    _, _, pi = rh.run_batch(c["edges"], c["kappa"], tg[:1], 2, True)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    lut = {int(l): i for i, l in enumerate(labels)}
    og = orc.OracleGraph(*gg.build_csr(len(labels), ne, c["kappa"]))
    checked = 0
    for a, b in tg[:10]:
        if int(a) not in lut or int(b) not in lut:
            continue
        st = rh.run_one_stages(pi, int(a), int(b), 2)
        if st["ncomp"] != 1 or st.get("PD1") is None:
            continue
        o = og.run_one(lut[int(a)], lut[int(b)], hop=2, flags=orc.F_NORM | orc.F_EXTENDED)
        assert list(o["vert"]) == st["nodes"]
        assert np.array_equal(o["fval"], np.array([st["fval"][x] for x in st["nodes"]]))
        k = o["pkind"]
        assert np.array_equal(np.stack([o["pbirth"], o["pdeath"]], 1)[k != orc.K_ONE], np.array(st["PD0"]).reshape(-1, 2))
        assert np.array_equal(np.stack([o["pbirth"], o["pdeath"]], 1)[k == orc.K_ONE], np.array(st["PD1"]).reshape(-1, 2))
        vert = o["vert"]
        assert np.stack([vert[o["elo"]][o["pos"]], vert[o["ehi"]][o["pos"]]], 1).tolist() == [list(e) for e in st["Pos"]]
        assert np.stack([vert[o["elo"]][o["neg"]], vert[o["ehi"]][o["neg"]]], 1).tolist() == [list(e) for e in st["Neg"]]
        checked += 1
    assert checked >= 3


def test_kd_generator_asis_vs_oracle():
    """Knowledge_Distillation/data_utils_NC.compute_persistence_image (filt='ricci', mode='PI')
    imported AS-IS (oracle/ref_harness.load_kd) on a fresh PPI-shaped graph vs the oracle's node mode + KD flags."""
    from helpers import sorted_rows
    c = gg.make_config("ppi", scale=0.06, continuous=True)
    edges, kappa = c["edges"], c["kappa"]
    g = rh.build_nx_graph(edges)
    ricci = rh.ricci_list(edges, [float(k) for k in kappa])
    labels, ne = gg.relabel_first_appearance(edges)
    csr = gg.build_csr(len(labels), ne, kappa)
    og = orc.OracleGraph(*csr)
    lut = {int(l): i for i, l in enumerate(labels)}
    fl = orc.F_NORM | orc.F_EXTENDED | orc.F_KEEP_ZERO | orc.F_NORM_EPS
    rng = np.random.default_rng(5)
    for u_old in rng.choice(labels, 6, replace=False):
        for hop in (1, 2):
            r = rh.kd_run_node(g, ricci, int(u_old), hop)
            o = og.run_one(lut[int(u_old)], lut[int(u_old)], hop=hop, mode=orc.MODE_NODE, flags=fl)
            if r is None:
                assert o["status"] == 2
                continue
            newid = np.array([lut[int(x)] for x in r["old_label"]])
            order = np.argsort(newid)
            assert np.array_equal(newid[order], o["vert"])
            assert np.array_equal(r["filt"][order], o["fval"])
            up, one = o["pkind"] == 0, o["pkind"] == 4
            assert np.array_equal(sorted_rows(r["ord0"]), sorted_rows(np.stack([o["pbirth"][up], o["pdeath"][up]], 1)))
            assert np.array_equal(sorted_rows(r["ext1"]), sorted_rows(np.stack([o["pbirth"][one], o["pdeath"][one]], 1)))
            assert rel_err(r["pi"], o["img"]) < 1e-10


@pytest.mark.parametrize("hop,descriptor,ext", [(3, "sum", True), (1, "max", True), (2, "min", False)])
def test_edge_cases_match_reference(hop, descriptor, ext):
    """targets the committed fixtures do not hold: self pairs (u, u), repeated targets, reversed pairs, unknown labels,
    far-apart pairs; hop 3; min / max descriptors with the loops -- the as-is reference batch call vs the oracle."""
    c = gg.make_config("pubmed", scale=0.04, continuous=True)
    e = c["edges"]
    rng = np.random.default_rng(23)
    pick = e[rng.choice(len(e), 10, replace=False)]
    un = np.unique(e)
    tg = np.concatenate([pick, pick[:3], pick[:3, ::-1], np.stack([un[:4], un[:4]], 1), rng.choice(un, size=(6, 2)),
                         np.array([[10 ** 7, int(un[0])], [int(un[1]), 10 ** 7 + 1]])])
    pi_ref, cnt_ref, _ = rh.run_batch(e, c["kappa"], tg, hop, ext, descriptor=descriptor)
    labels, ne = gg.relabel_first_appearance(e)
    lut = {int(l): i for i, l in enumerate(labels)}
    new_t = np.array([[lut.get(int(a), -1), lut.get(int(b), -1)] for a, b in tg], dtype=np.int32)
    og = orc.OracleGraph(*gg.build_csr(len(labels), ne, c["kappa"]))
    r = og.run_batch(new_t, hop=hop, descriptor=descriptor, flags=orc.F_NORM | (orc.F_EXTENDED if ext else 0))
    assert r["cnt_compute"] == cnt_ref
    assert np.array_equal(pi_ref.any(axis=1), r["pi"].any(axis=1))
    assert rel_err(r["pi"], pi_ref) < 1e-11
    assert np.array_equal(r["pi"][:3], r["pi"][10:13])                       # repeated targets: identical rows
