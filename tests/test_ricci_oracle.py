"""SURVEY.md row N4 -- the Ollivier-Ricci (Sinkhorn) curvature precompute.  PARITY UNPINNED by the reference (its two
third-party packages are neither vendored, pinned nor installed here): the CPU restatement oracle/ricci_oracle.py is
pinned against closed forms of the exact transport problem on toy graphs, to the accuracy the entropic regularisation
(reg = 0.1 against hop costs) leaves."""
import numpy as np
import pytest

import ricci_oracle as ro
from tlc_b200 import graphgen as gg


def csr_of(n, edges):
    e = np.asarray(edges, dtype=np.int64)
    return gg.build_csr(n, e, np.zeros(len(e)))[:2]


def cycle(n):
    return csr_of(n, [(i, (i + 1) % n) for i in range(n)])


def complete(n):
    return csr_of(n, [(i, j) for i in range(n) for j in range(i + 1, n)])


def test_long_cycle_is_flat():
    rp, col = cycle(12)
    k = ro.compute_ricci_curvature(rp, col)
    assert np.allclose(k, 0.0, atol=1e-2)          # W1 = 1 on a cycle of length >= 6: kappa = 0


def test_complete_graph_closed_form():
    for n in (4, 5, 8):
        rp, col = complete(n)
        k = ro.compute_ricci_curvature(rp, col)
        expect = 1.0 - abs(0.5 - 0.5 / (n - 1))    # move the excess alpha - (1-alpha)/(n-1) from x to y
        assert np.allclose(k, expect, atol=1e-2), (n, k[:3], expect)


def test_star_and_path_against_the_exact_lp():
    # leaf - centre edges of a star, and the edges of a short path: exact W1 from a linear programme
    star = csr_of(6, [(0, i) for i in range(1, 6)])
    path = csr_of(6, [(i, i + 1) for i in range(5)])
    for rp, col in (star, path):
        for x in range(len(rp) - 1):
            for e in range(rp[x], rp[x + 1]):
                y = int(col[e])
                a, src = ro.support(rp, col, x)
                b, tgt = ro.support(rp, col, y)
                d = ro.hop_costs(rp, col, src, tgt)
                assert abs(ro.sinkhorn2(a, b, d)[0] - ro.exact_w1(a, b, d)) < 1e-2
                assert abs(ro.edge_curvature(rp, col, x, y) - ro.edge_curvature(rp, col, y, x)) < 1e-9   # symmetric


def test_random_graph_against_the_exact_lp_and_mass_conservation():
    c = gg.make_config("cora", scale=0.05)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    rp, col, _ = gg.build_csr(len(labels), ne, c["kappa"])
    rng = np.random.default_rng(0)
    for x, y in ne[rng.choice(len(ne), 12, replace=False)]:
        a, src = ro.support(rp, col, int(x))
        b, tgt = ro.support(rp, col, int(y))
        assert abs(a.sum() - 1.0) < 1e-12 and abs(b.sum() - 1.0) < 1e-12
        d = ro.hop_costs(rp, col, src, tgt)
        assert d.max() <= 3 and d.min() >= 0           # supports of an edge are at most 3 hops apart
        m, it = ro.sinkhorn2(a, b, d)
        assert it <= ro.NUM_ITER_MAX
        assert abs(m - ro.exact_w1(a, b, d)) < 1e-2
        assert -2.0 <= 1.0 - m <= 1.0
