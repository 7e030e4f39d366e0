"""Kernel S (k0_small.cu, the fused small-vicinity kernels) through the C-ABI vs the oracle: pairs bit-exact
(kind, birth / death vertex, float64 birth / death value, in the reference's order), images 1e-5, statuses equal;
and the batch call, which runs kernel S first and hands larger vicinities to the staged pipeline, vs the oracle and
vs the staged pipeline alone."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as orc
from helpers import GRAPH_CASES, load_case, rel_err
from tlc_b200 import _lib as L
from tlc_b200 import api, graphgen as gg

IMG_TOL = 1e-5  # north_star: persistence images within 1e-5 relative


def check_diagrams(g, og, targets, hop, descriptor, flags, oflags, mode=L.MODE_EDGE, img_mask=None, oimg_mask=None):
    """every target kernel S takes: status, sizes, pair sequence, image equal to the oracle's; returns #taken"""
    res = g.small_diagrams(targets, hop=hop, mode=mode, descriptor=descriptor, flags=flags, img_mask=img_mask)
    taken = 0
    for i, (u, v) in enumerate(targets):
        a = res[i]
        o = og.run_one(int(u), int(v), hop=hop, mode=mode, descriptor=descriptor, flags=oflags, img_mask=oimg_mask)
        ctx = "target %d (%d,%d) status gpu %d oracle %d n %d m %d" % (i, u, v, a["status"], o["status"], o["n"], o["m"])
        if a["status"] == L.ST_NOT_SMALL:
            assert o["n"] > 1024 or o["m"] > 4096, ctx  # beyond class C (class A: n <= 64, m <= 256; B: 256 / 2048; C: 1024 / 4096)
            continue
        taken += 1
        assert a["status"] == o["status"], ctx
        if o["status"] == 5:
            continue
        assert a["n"] == o["n"], ctx
        if o["status"] == 2 and o["n"] == 0:
            continue
        assert a["m"] == o["m"], ctx
        if o["status"] > 1 and o["status"] != 7:
            assert not a["img"].any(), ctx
            continue
        k = len(o["pkind"])
        assert len(a["pkind"]) == k, ctx
        assert np.array_equal(a["pkind"], o["pkind"]), ctx
        assert np.array_equal(a["pbv"], o["pbv"]) and np.array_equal(a["pdv"], o["pdv"]), ctx
        assert np.array_equal(a["pbirth"], o["pbirth"]) and np.array_equal(a["pdeath"], o["pdeath"]), ctx
        assert rel_err(a["img"], o["img"]) < IMG_TOL, ctx
    return taken


def make(name, scale=1.0, cont=False):
    c = gg.make_config(name, scale=scale, continuous=cont)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    csr = gg.build_csr(len(labels), ne, c["kappa"])
    return csr, ne, len(labels)


@pytest.mark.parametrize("tag", GRAPH_CASES)
@pytest.mark.parametrize("ext", [0, 1])
def test_small_golden_cases(tag, ext):
    c = load_case(tag)
    g = api.VicinityGraph(*c["csr"], device=0)
    og = orc.OracleGraph(*c["csr"])
    oflags = orc.F_NORM | (orc.F_EXTENDED if ext else 0)
    flags = L.F_NORM | (L.F_EXTENDED if ext else 0)
    check_diagrams(g, og, c["new_targets"], c["hop"], c["descriptor"], flags, oflags)
    # the batch call (kernel S + staged pipeline) against the REAL reference's pi_sg stored in the fixture
    pi, status, cnt = g.vicinity_pi(c["new_targets"], hop=c["hop"], descriptor=c["descriptor"], flags=flags)
    ref = c["pi_ext%d" % ext]
    assert cnt == int(c["cnt_ext%d" % ext])
    assert np.array_equal(ref.any(axis=1), pi.any(axis=1))
    assert rel_err(pi, ref) < IMG_TOL
    g.close()


@pytest.mark.parametrize("ext", [0, 1])
@pytest.mark.parametrize("name,scale,hop,cont", [("cora", 1.0, 2, False), ("cora", 1.0, 3, False), ("pubmed", 0.3, 2, False),
                                                 ("pubmed", 0.3, 2, True), ("pubmed", 1.0, 2, True), ("computers", 0.05, 2, True),
                                                 ("computers", 1.0, 1, False), ("computers", 1.0, 1, True), ("ppi", 0.3, 1, False),
                                                 ("collab", 0.05, 2, True)])
def test_small_random_targets(name, scale, hop, cont, ext):
    csr, ne, N = make(name, scale, cont)
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    rng = np.random.default_rng(11)
    tg = ne[rng.choice(len(ne), 160, replace=False)].astype(np.int32)
    neg = rng.integers(0, N, size=(32, 2)).astype(np.int32)
    tg = np.concatenate([tg, neg, np.array([[-1, 3], [0, 0], [5, N + 7]], np.int32)])
    oflags = orc.F_NORM | (orc.F_EXTENDED if ext else 0)
    flags = L.F_NORM | (L.F_EXTENDED if ext else 0)
    taken = check_diagrams(g, og, tg, hop, "sum", flags, oflags)
    assert taken > 0
    # batch call: kernel S first, the rest staged -- and the staged pipeline alone -- both equal to the oracle
    o = og.run_batch(tg, hop=hop, flags=oflags)
    for fl in (flags, flags | L.F_NO_SMALL):
        pi, status, cnt = g.vicinity_pi(tg, hop=hop, flags=fl)
        assert np.array_equal(status, o["status"])
        assert cnt == o["cnt_compute"]
        assert rel_err(pi, o["pi"]) < IMG_TOL
        if not (fl & L.F_NO_SMALL):
            s = g.last_small()
            assert s["rows_a"] + s["rows_b"] + s["rows_c"] + s["rows_staged"] == len(tg)
            assert s["rows_a"] + s["rows_b"] + s["rows_c"] > 0
    g.close()


@pytest.mark.parametrize("descriptor", ["min", "max", "sum"])
@pytest.mark.parametrize("norm", [0, 1])
def test_small_descriptors(descriptor, norm):
    csr, ne, N = make("pubmed", 0.3, True)
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    tg = ne[np.random.default_rng(3).choice(len(ne), 96, replace=False)].astype(np.int32)
    check_diagrams(g, og, tg, 2, descriptor, (L.F_NORM if norm else 0) | L.F_EXTENDED, (orc.F_NORM if norm else 0) | orc.F_EXTENDED)
    g.close()


@pytest.mark.parametrize("plain", [0, 1])
def test_small_node_mode_kd_flags(plain):
    """PDGNN generator flags: node-centred ball, zero-persistence pairs kept, eps-normalisation, Ord0 u Ext1 image"""
    csr, ne, N = make("pubmed", 0.3, True)
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    ids = np.random.default_rng(5).choice(N, 128, replace=False).astype(np.int32)
    tg = np.stack([ids, ids], 1)
    fl = L.F_NORM | L.F_EXTENDED | L.F_KEEP_ZERO | L.F_NORM_EPS | (L.F_SUM_PLAIN if plain else 0)
    ofl = orc.F_NORM | orc.F_EXTENDED | orc.F_KEEP_ZERO | orc.F_NORM_EPS | (orc.F_SUM_PLAIN if plain else 0)
    taken = check_diagrams(g, og, tg, 2, "sum", fl, ofl, mode=L.MODE_NODE)
    assert taken > 0
    o = og.run_batch(tg, hop=2, mode=orc.MODE_NODE, flags=ofl)
    pi, status, cnt = g.vicinity_pi(tg, hop=2, mode=L.MODE_NODE, flags=fl)
    assert np.array_equal(status, o["status"]) and cnt == o["cnt_compute"]
    assert rel_err(pi, o["pi"]) < IMG_TOL
    g.close()


def test_small_forced_roots_mode():
    """the PDGNN link-prediction generator's vicinity (roots always members, data_utils_LP.py:107-118)"""
    csr, ne, N = make("pubmed", 0.3, True)
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    rng = np.random.default_rng(9)
    tg = np.concatenate([ne[rng.choice(len(ne), 64, replace=False)], rng.integers(0, N, size=(64, 2))]).astype(np.int32)
    fl = L.F_NORM | L.F_EXTENDED | L.F_KEEP_ZERO | L.F_NORM_EPS
    ofl = orc.F_NORM | orc.F_EXTENDED | orc.F_KEEP_ZERO | orc.F_NORM_EPS
    res = g.small_diagrams(tg, hop=2, mode=L.MODE_EDGE_FORCED, flags=fl)
    for i, (u, v) in enumerate(tg):
        o = og.run_one(int(u), int(v), hop=2, mode=orc.MODE_EDGE_FORCED, flags=ofl)
        a = res[i]
        if a["status"] == L.ST_NOT_SMALL:
            continue
        # disconnected forced vicinities: the reference has no connectivity check there -> status 3 on both sides
        assert a["status"] == o["status"], (i, u, v, a["status"], o["status"])
        if o["status"] <= 1:
            assert np.array_equal(a["pbirth"], o["pbirth"]) and np.array_equal(a["pdeath"], o["pdeath"])
            assert np.array_equal(a["pbv"], o["pbv"]) and np.array_equal(a["pdv"], o["pdv"])
            assert rel_err(a["img"], o["img"]) < IMG_TOL
    g.close()


def test_small_near_ties_and_exact_ties():
    """hop-distance filtration (kappa = 0: hundreds of exact ties) and values 3e-7 apart (perturbed-key crossings, F4)"""
    csr, ne, N = make("cora", 1.0, False)  # kappa == 0
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    tg = ne[np.random.default_rng(1).choice(len(ne), 256, replace=False)].astype(np.int32)
    check_diagrams(g, og, tg, 2, "sum", L.F_NORM | L.F_EXTENDED, orc.F_NORM | orc.F_EXTENDED)
    g.close()
    rowptr, col, kappa = csr
    rng = np.random.default_rng(2)
    # curvature levels 3e-7 apart: distinct vertex values closer than the 1e-6 perturbation
    lev = rng.integers(0, 4, size=len(ne)) * 3e-7
    c2 = gg.build_csr(N, ne, lev)
    g = api.VicinityGraph(*c2, device=0)
    og = orc.OracleGraph(*c2)
    check_diagrams(g, og, tg, 2, "sum", L.F_EXTENDED, orc.F_EXTENDED)          # un-normalised: values stay 3e-7 apart
    check_diagrams(g, og, tg, 2, "sum", L.F_NORM | L.F_EXTENDED, orc.F_NORM | orc.F_EXTENDED)
    g.close()


def test_small_full_cora_batch_equals_staged():
    """every edge of the Cora-shaped graph: kernel S rows == staged rows bit for bit (float64)"""
    csr, ne, N = make("cora")
    g = api.VicinityGraph(*csr, device=0)
    tg = ne.astype(np.int32)
    for ext in (0, 1):
        fl = L.F_NORM | (L.F_EXTENDED if ext else 0)
        pi_s, st_s, cnt_s = g.vicinity_pi(tg, hop=2, flags=fl)
        s = g.last_small()
        assert s["rows_a"] + s["rows_b"] + s["rows_c"] + s["rows_staged"] == len(tg)
        assert s["rows_staged"] == 0  # (the largest Cora-shaped 2-hop vicinity has a few hundred vertices: class C)
        pi_g, st_g, cnt_g = g.vicinity_pi(tg, hop=2, flags=fl | L.F_NO_SMALL)
        assert np.array_equal(st_s, st_g) and cnt_s == cnt_g
        assert rel_err(pi_s, pi_g) < 1e-12
    g.close()


@pytest.mark.parametrize("mode,omode", [(L.MODE_EDGE_UNION, orc.MODE_EDGE_UNION), (L.MODE_EDGE_REMOVEINTER, orc.MODE_EDGE_REMOVEINTER)])
def test_union_and_removeinter_ranges(mode, omode):
    """the vicinity shapes of the legacy sg2pimg(range='union' / 'removeinter'), riccidist2dgm.py:242-247, 289-296:
    kernel S and the staged pipeline against the oracle"""
    csr, ne, N = make("pubmed", 0.3, True)
    g = api.VicinityGraph(*csr, device=0)
    og = orc.OracleGraph(*csr)
    rng = np.random.default_rng(21)
    tg = np.concatenate([ne[rng.choice(len(ne), 96, replace=False)], rng.integers(0, N, size=(16, 2))]).astype(np.int32)
    fl, ofl = L.F_NORM | L.F_EXTENDED, orc.F_NORM | orc.F_EXTENDED
    taken = check_diagrams(g, og, tg, 1, "sum", fl, ofl, mode=mode)
    assert taken > 0
    o = og.run_batch(tg, hop=1, mode=omode, flags=ofl)
    for f in (fl, fl | L.F_NO_SMALL):
        pi, status, cnt = g.vicinity_pi(tg, hop=1, mode=mode, flags=f)
        assert np.array_equal(status, o["status"]) and cnt == o["cnt_compute"]
        assert rel_err(pi, o["pi"]) < IMG_TOL
    g.close()


@pytest.mark.parametrize("t", [0.1, 1.0, 5.0])
def test_hks_filtration(t):
    """filt='hks' of the PDGNN generators (data_utils_NC.py:87-93,115-117): the kernel's Taylor evaluation against the
    reference's own numpy / scipy lines (eigh of the normalised Laplacian) to 1e-10, and everything downstream of the
    filtration bit-exact against the oracle fed with the SAME values"""
    csr, ne, N = make("pubmed", 0.3, True)
    g = api.VicinityGraph(*csr, device=0)
    g.set_hks_time(t)
    og = orc.OracleGraph(*csr)
    ids = np.random.default_rng(5).choice(N, 24, replace=False).astype(np.int32)
    tg = np.stack([ids, ids], 1)
    fl = L.F_NORM | L.F_EXTENDED | L.F_KEEP_ZERO | L.F_NORM_EPS | L.F_FILT_HKS
    ofl = orc.F_NORM | orc.F_EXTENDED | orc.F_KEEP_ZERO | orc.F_NORM_EPS
    d = g.vicinity_detail(tg, hop=2, mode=L.MODE_NODE, flags=fl)
    checked = 0
    for i, (u, _) in enumerate(tg):
        a = g.per_target(d, i)
        if a["status"] > 1:
            continue
        ref = orc.hks_signature(a["n"], a["elo"], a["ehi"], time=t)
        assert np.max(np.abs(a["fval"] - ref)) < 1e-10, (i, np.max(np.abs(a["fval"] - ref)))
        o = og.run_one(int(u), int(u), hop=2, mode=orc.MODE_NODE, flags=ofl, fval=a["fval"])
        assert a["status"] == o["status"]
        assert np.array_equal(a["pkind"], o["pkind"]) and np.array_equal(a["pbv"], o["pbv"]) and np.array_equal(a["pdv"], o["pdv"])
        assert np.array_equal(a["pbirth"], o["pbirth"]) and np.array_equal(a["pdeath"], o["pdeath"])
        assert rel_err(a["img"], o["img"]) < IMG_TOL
        checked += 1
    assert checked >= 12
    g.close()
