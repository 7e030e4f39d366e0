"""The drop-in `sg2dgm` mirror against the REAL reference's stored outputs (golden fixtures) and against
its error behaviour; reads like a test the reference could have had (it ships none)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import networkx as nx

from helpers import GRAPH_CASES, load_case, rel_err, seg


def build(c):
    import sg2dgm.riccidist2dgm as r
    g = nx.Graph()
    g.add_edges_from([(int(a), int(b)) for a, b in c["edges"]])           # loaddatas.py:88-92
    ricci = sorted([[int(a), int(b), float(k)] for (a, b), k in zip(c["edges"], c["kappa"])] +
                   [[int(b), int(a), float(k)] for (a, b), k in zip(c["edges"], c["kappa"])])  # :117-121
    return r.graph2pi(g, ricci_curv=ricci)


@pytest.mark.parametrize("tag", GRAPH_CASES)
def test_get_pimg_for_all_edges_matches_reference(tag):
    c = load_case(tag)
    pi = build(c)
    for ext in (False, True):
        pi.get_pimg_for_all_edges([list(map(int, t)) for t in c["targets"]], cores=16, hop=c["hop"], norm=True,
                                  extended_flag=ext, resolution=5, descriptor=c["descriptor"])   # loaddatas.py:100
        ref = c["pi_ext%d" % ext]
        assert pi.pi_sg.shape == ref.shape and pi.pi_sg.dtype == np.float64
        assert pi.cnt_compute == int(c["cnt_ext%d" % ext])
        assert rel_err(pi.pi_sg, ref) < 1e-5


def test_error_behaviour_toy():
    c = load_case("toy7_hop2")
    pi = build(c)
    dn = pi.dict_node
    with pytest.raises(AssertionError):
        pi.sg2dgm_accelerate(dn[0], dn[5], 2, descriptor="sum", norm=True)      # empty vicinity  :318
    with pytest.raises(KeyError):
        pi.sg2dgm_accelerate(dn[1], dn[2], 2, norm=True)                        # default descriptor "seal"
    with pytest.raises(ZeroDivisionError):
        pi.sg2dgm_accelerate(dn[4], dn[5], 1, descriptor="sum", norm=True)      # vicinity == {u,v}
    img = pi.sg2dgm_accelerate(dn[1], dn[2], 2, descriptor="sum", norm=True, extended_flag=True)
    assert img.shape == (5, 5) and abs(img[0, 0] - 0.00780321722227386) < 1e-12
    z = pi.get_pimg_for_one_edge(0, 99, hop=2, descriptor="sum")                # unknown label -> zeros, no raise
    assert z.shape == (25,) and not z.any()


def test_union_find_and_imager_mirrors():
    import sg2dgm.accelerated_PD as apd
    import sg2dgm.PersistenceImager as pimg
    c = load_case("pubmed_s_hop2_dyadic")
    for k in range(5):
        if int(c["stage_ncomp"][k]) != 1:
            continue
        nodes = seg(c, "nodes", k)
        fval = seg(c, "fval", k)
        neg = seg(c, "neg", k).reshape(-1, 2)
        pos = seg(c, "pos", k).reshape(-1, 2)
        # rebuild the canonical vicinity graph the fixture was generated from
        h = nx.Graph()
        h.add_nodes_from(int(x) for x in nodes)
        es = sorted([tuple(sorted(map(int, e))) for e in np.concatenate([neg, pos])])
        h.add_edges_from(es)
        for x, f in zip(nodes, fval):
            h.nodes[int(x)]["sum"] = float(f)
        sf = apd.perturb_filter_function(h, "sum")
        PD, Pos, Neg = apd.Union_find(sf)
        assert np.array_equal(np.array(PD), seg(c, "pd0", k).reshape(-1, 2))
        assert Pos == pos.tolist() and Neg == neg.tolist()
        PD1 = apd.Accelerate_PD(Pos, Neg, sf)
        assert np.array_equal(np.array(PD1).reshape(-1, 2), seg(c, "pd1", k).reshape(-1, 2))
        img = pimg.PersistenceImager(resolution=5).transform(np.array(PD + PD1))
        assert img.shape == (5, 5)


# ---- PDGNN generator mirror (sg2dgm/kd.py) vs the UNMODIFIED data_utils_NC.compute_persistence_image outputs ----
@pytest.mark.parametrize("tag", __import__("helpers").KD_CASES)
def test_kd_generator_tuple_matches_reference(tag):
    from helpers import kd_expected, load_kd_case, sorted_rows
    import sg2dgm.kd as kd
    c = load_kd_case(tag)
    pi = build(c)
    res = kd.compute_persistence_images(pi, [int(u) for u in c["nodes"]], hop=c["hop"], filt=c["filt_name"])
    assert len(res) == len(c["nodes"])
    for k, r in enumerate(res):
        e = kd_expected(c, k)
        if e["none"]:
            assert len(r) == 2 and r[0] is None and r[1] is None          # data_utils_NC.py:103-104
            continue
        assert len(r) == 9
        ord0, ext1, img, filt, edge_index, pi0, pi1, _, _ = r
        newid = np.array([c["lut"][int(x)] for x in r.old_label])        # graph ids in first-appearance numbering
        assert np.array_equal(newid, e["vert"])                           # canonical local order
        assert np.array_equal(np.asarray(filt), e["filt"])                # bit-exact float64
        eg = newid[np.asarray(edge_index)].T
        assert np.array_equal(eg, e["edges"])                             # induced edge set, lexicographic
        assert ord0.shape == (len(e["vert"]) - 1, 2) and ext1.shape == (len(e["edges"]) - len(e["vert"]) + 1, 2)
        assert np.array_equal(sorted_rows(ord0), e["ord0"])               # bit-exact multisets
        assert np.array_equal(sorted_rows(ext1), e["ext1"])
        assert rel_err(img, e["pi"]) < 1e-5 and rel_err(pi0, e["pi0"]) < 1e-5 and rel_err(pi1, e["pi1"]) < 1e-5
    # per-node signature of the reference (data_utils_NC.py:95)
    u0 = int(c["nodes"][0])
    one = kd.compute_persistence_image(pi, u0, filt=c["filt_name"], hop=c["hop"], mode="PI")
    if res[0][0] is None:
        assert one[0] is None and one[1] is None
    else:
        # (the image sums its points in CTA-size dependent order: last-bit differences between differently sized calls)
        assert np.array_equal(one[0], res[0][0]) and rel_err(one[2], res[0][2]) < 1e-12


@pytest.mark.parametrize("tag", __import__("helpers").KD_LP_CASES)
def test_kd_lp_generator_tuple_matches_reference(tag):
    """edge-centred PDGNN generator mirror vs the UNMODIFIED data_utils_LP.compute_persistence_image outputs."""
    from helpers import kd_expected, load_kd_case, sorted_rows
    import sg2dgm.kd as kd
    c = load_kd_case(tag)
    pi = build(c)
    res = kd.compute_persistence_images_lp(pi, [(int(u), int(v)) for u, v in c["nodes"]], hop=c["hop"])
    ok = 0
    for k, r in enumerate(res):
        e = kd_expected(c, k)
        if e["none"] or int(c["kind"][k]) == 2 or r[0] is None:
            assert r[0] is None and r[1] is None      # `return None, None` / outside the contract (disconnected)
            continue
        ord0, ext1, img, filt, edge_index, pi0, pi1, _, _ = r
        newid = np.array([c["lut"][int(x)] for x in r.old_label])
        assert np.array_equal(newid, e["vert"]) and np.array_equal(np.asarray(filt), e["filt"])
        assert np.array_equal(newid[np.asarray(edge_index)].T, e["edges"])
        assert np.array_equal(sorted_rows(ord0), e["ord0"]) and np.array_equal(sorted_rows(ext1), e["ext1"])
        assert rel_err(img, e["pi"]) < 1e-5 and rel_err(pi0, e["pi0"]) < 1e-5 and rel_err(pi1, e["pi1"]) < 1e-5
        ok += 1
    assert ok >= 10
    u0, v0 = map(int, c["nodes"][0])
    one = kd.compute_persistence_image(pi, u0, v0, filt="ricci", hop=c["hop"], mode="PI")
    assert (one[0] is None) == (res[0][0] is None)


def test_kd_gc_generator_matches_reference():
    """graph-classification PDGNN generator mirror vs the UNMODIFIED data_utils_GC.compute_persistence_image (filt='degree')."""
    from helpers import load_kd_gc_case, sorted_rows
    import sg2dgm.kd as kd
    c = load_kd_gc_case()
    res = kd.compute_persistence_images_gc(c["graphs"], filt=c["filt_name"])
    fo, o0, o1 = c["kd_filt_off"], c["kd_ord0_off"], c["kd_ext1_off"]
    for k, r in enumerate(res):
        if c["none"][k]:
            assert r[0] is None and r[1] is None                               # data_utils_GC.py:99-100
            continue
        ord0, ext1, img, filt, edge_index, pi0, pi1, _, _ = r
        assert np.array_equal(np.asarray(r.old_label), np.arange(c["sizes"][k]))
        assert np.array_equal(np.asarray(filt), c["kd_filt"][fo[k]:fo[k + 1]])
        n, e = c["graphs"][k]
        assert np.array_equal(np.asarray(edge_index).T, np.unique(np.sort(e, axis=1), axis=0))
        assert np.array_equal(sorted_rows(ord0), sorted_rows(c["kd_ord0"][2 * o0[k]:2 * o0[k + 1]]))
        assert np.array_equal(sorted_rows(ext1), sorted_rows(c["kd_ext1"][2 * o1[k]:2 * o1[k + 1]]))
        assert rel_err(img, c["pi"][k]) < 1e-5 and rel_err(pi0, c["pi0"][k]) < 1e-5 and rel_err(pi1, c["pi1"][k]) < 1e-5


def test_kd_emitter_batches_large_requests():
    """the PDGNN emitter splits a long node list into several C-ABI calls (every intermediate comes back to the host):
    a tiny budget forces many batches; results equal those of one call."""
    import sg2dgm.kd as kd
    import sg2dgm.riccidist2dgm as r
    from tlc_b200 import graphgen as gg
    c = gg.make_config("ppi", scale=0.2)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    pi = r.graph2pi.from_csr(*gg.build_csr(len(labels), ne, c["kappa"]))
    nodes = list(range(0, 120))
    one = kd.compute_persistence_images(pi, nodes, hop=1)
    many = kd.compute_persistence_images(pi, nodes, hop=1, budget=2000)
    assert len(one) == len(many) == len(nodes)
    for a, b in zip(one, many):
        assert (a[0] is None) == (b[0] is None)
        if a[0] is None:
            continue
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(np.asarray(a[3]), np.asarray(b[3]))
        assert rel_err(a[2], b[2]) < 1e-12 and np.array_equal(a[4], b[4])
