"""SURVEY.md row N4 on the GPU: kernel 6 (tlc_ollivier_ricci) against the CPU restatement oracle/ricci_oracle.py of the
published GraphRicciCurvature + POT algorithm (parity unpinned by the reference, see that file), and against closed forms
of the exact transport problem on toy graphs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ricci_oracle as ro
from tlc_b200 import api, graphgen as gg

TOL = 1e-9  # float64 Sinkhorn on both sides; only the summation order inside the matrix-vector products differs


def csr_of(n, edges):
    e = np.asarray(edges, dtype=np.int64)
    return gg.build_csr(n, e, np.zeros(len(e)))[:2]


def test_toy_graphs_closed_forms():
    for n in (4, 5, 8):
        rp, col = csr_of(n, [(i, j) for i in range(n) for j in range(i + 1, n)])
        k = api.ollivier_ricci(rp, col)
        assert np.allclose(k, 1.0 - abs(0.5 - 0.5 / (n - 1)), atol=1e-2)
    rp, col = csr_of(12, [(i, (i + 1) % 12) for i in range(12)])
    assert np.allclose(api.ollivier_ricci(rp, col), 0.0, atol=1e-2)


@pytest.mark.parametrize("name,scale", [("cora", 0.2), ("pubmed", 0.05), ("computers", 0.03)])
def test_random_graph_matches_restatement(name, scale):
    c = gg.make_config(name, scale=scale)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    rp, col, _ = gg.build_csr(len(labels), ne, c["kappa"])
    k, it = api.ollivier_ricci(rp, col, return_iters=True)
    rng = np.random.default_rng(1)
    # both directions of an edge carry the same value
    for x in rng.choice(len(labels), 20):
        for q in range(rp[x], rp[x + 1]):
            y = col[q]
            m = rp[y] + np.searchsorted(col[rp[y]:rp[y + 1]], x)
            assert k[q] == k[m]
    for x, y in ne[rng.choice(len(ne), 40, replace=False)]:
        q = rp[x] + np.searchsorted(col[rp[x]:rp[x + 1]], y)
        a, src = ro.support(rp, col, int(min(x, y)))
        b, tgt = ro.support(rp, col, int(max(x, y)))
        d = ro.hop_costs(rp, col, src, tgt)
        m, cpt = ro.sinkhorn2(a, b, d)
        assert abs(k[q] - (1.0 - m)) < TOL, (x, y, k[q], 1.0 - m)
        assert it[q] == cpt
    assert np.all(k <= 1.0 + 1e-12) and np.all(k >= -2.0)


def test_mirror_returns_the_reference_layout():
    from sg2dgm.feature_cache import compute_ricci_curvature
    ei = np.array([[10, 20, 30, 10, 20, 10], [20, 30, 10, 40, 10, 10]])   # a duplicate (20,10) and a self-loop (10,10)
    rl = compute_ricci_curvature(edge_index=ei)
    assert rl == sorted(rl) and len(rl) == 8                                   # 4 undirected edges, both directions
    d = {(a, b): k for a, b, k in rl}
    assert all(d[(a, b)] == d[(b, a)] for a, b in d)
