"""The claim behind kernel 2v's no-pair certificate (k2v_vorder.cu step 0), checked on the CPU against the
oracle's restatement of the reference sweep (accelerated_PD.py:40-68) on random graphs with caller-supplied filtration values:
whenever the certificate holds -- values in [0, 1], adjacent roots at 0, every other vertex either has a neighbour more than
delta lower (A) or sits on a plateau of exactly equal values with no near-tie neighbour that is grounded in an A vertex (B) --
the ascending sweep emits NO ordinary pair and the vicinity is connected.  Plateaus, exact ties, near-ties (1e-10 .. 1e-16
apart: the perturbed-key crossings of SURVEY.md F4) and extra local minima are all generated; the test also checks that the
certificate is not vacuous and that it does fail where pairs exist."""
import numpy as np

import oracle as orc
from tlc_b200 import graphgen as gg

DELTA = 1e-8


def certificate(n, rows, f, lu, lv):
    """numpy restatement of the kernel's test; rows[x] = neighbour ids of x"""
    if f.max() > 1.0 or f[lu] != 0.0 or f[lv] != 0.0:
        return False
    if lu != lv and lv not in rows[lu]:
        return False
    cls = np.zeros(n, np.int8)   # 1: A or grounded B, 2: B not grounded yet
    cls[lu] = cls[lv] = 1
    for x in range(n):
        if x in (lu, lv):
            continue
        if not f[x] > 0.0:
            return False
        fy = f[rows[x]]
        if (fy < f[x] - DELTA).any():
            cls[x] = 1
        elif ((fy >= f[x] - DELTA) & (fy < f[x])).any() or ((fy > f[x]) & (fy <= f[x] + DELTA)).any() or not (fy == f[x]).any():
            return False
        else:
            cls[x] = 2
    while (cls == 2).any():
        ch = False
        for x in np.flatnonzero(cls == 2):
            nb = rows[x]
            if ((f[nb] == f[x]) & (cls[nb] == 1)).any():
                cls[x] = 1
                ch = True
        if not ch:
            return False
    return True


def random_case(rng):
    n = int(rng.integers(4, 28))
    # a connected graph: random tree + extra edges; vertex 0 - 1 is the target edge
    edges = {(0, 1)}
    for x in range(2, n):
        y = int(rng.integers(0, x))
        edges.add((y, x))
    for _ in range(int(rng.integers(0, 2 * n))):
        a, b = sorted(int(q) for q in rng.integers(0, n, 2))
        if a != b:
            edges.add((a, b))
    e = np.array(sorted(edges), dtype=np.int64)
    kind = rng.integers(0, 4)
    if kind == 0:      # hop-distance-like levels: many exact ties and plateaus
        f = rng.integers(1, 5, n).astype(np.float64) / 4.0
    elif kind == 1:    # BFS levels from the roots (the shape of a distance-to-roots filtration) + plateaus
        rows = [[] for _ in range(n)]
        for a, b in e:
            rows[a].append(b); rows[b].append(a)
        d = np.full(n, -1); d[0] = d[1] = 0
        q = [0, 1]
        while q:
            x = q.pop(0)
            for y in rows[x]:
                if d[y] < 0:
                    d[y] = d[x] + 1; q.append(y)
        f = d / max(1, d.max())
    elif kind == 2:    # continuous values
        f = rng.random(n)
    else:              # levels with near-ties sprinkled in
        f = rng.integers(1, 4, n).astype(np.float64) / 4.0 + rng.choice([0.0, 1e-10, 1e-13, 2.0 ** -52, -1e-12], n)
    f = np.clip(f, 1e-3, 1.0)
    f[0] = f[1] = 0.0
    f[int(rng.integers(2, n))] = 1.0
    return n, e, f


def test_certificate_implies_no_ordinary_pair():
    rng = np.random.default_rng(20260)
    held = failed_with_pairs = 0
    for _ in range(1500):
        n, e, f = random_case(rng)
        rp, col, kap = gg.build_csr(n, e, np.zeros(len(e)))
        rows = [col[rp[x]:rp[x + 1]] for x in range(n)]
        og = orc.OracleGraph(rp, col, kap)
        a = og.run_one(0, 1, hop=n, flags=0, fval=f)      # hop >= diameter: the vicinity is the whole graph, ids = local ids
        assert a["n"] == n and a["lu"] == 0 and a["lv"] == 1
        n_up = int(np.count_nonzero(a["pkind"] == orc.K_UP))
        ok = certificate(n, rows, f, 0, 1)
        if ok:
            held += 1
            assert a["status"] == 0 and n_up == 0, (n, e.tolist(), f.tolist(), a["pbirth"], a["pdeath"])
            ess = np.flatnonzero(a["pkind"] == orc.K_ESS)
            assert len(ess) == 1 and a["pbirth"][ess[0]] == 0.0 and a["pdeath"][ess[0]] == f.max()
            assert a["pbv"][ess[0]] == 0 and a["pdv"][ess[0]] == int(np.flatnonzero(f == f.max())[0])
        elif n_up > 0:
            failed_with_pairs += 1
    assert held > 200 and failed_with_pairs > 50, (held, failed_with_pairs)   # neither side of the claim is vacuous


def test_near_tie_above_a_plateau_is_rejected():
    """why class B excludes near-tie neighbours: a plateau vertex x (value F) whose neighbour z is ONE ulp higher, with F just
    below a binade boundary so that fl(F + t) == fl(f_z + t): the edge z-x ties with the plateau's own key, precedes it in
    the canonical order, and the reference emits the pair [F, f_z] of persistence 5.5e-17.  The certificate must say no."""
    F = None
    for k in range(1, 4000):
        c = 0.5 - k * 1e-9
        t = (c + 1.0) * 1e-6
        if c + t == np.nextafter(c, 1.0) + t and c + t >= 0.5:
            F = c
            break
    assert F is not None
    fz = float(np.nextafter(F, 1.0))
    # ids: 0, 1 roots; x = 2 (F), z = 3 (fz, also next to a root), m = 4 (F, next to a root), 5 carries the maximum
    e = np.array([(0, 1), (0, 3), (0, 4), (2, 3), (2, 4), (0, 5)], dtype=np.int64)
    f = np.array([0.0, 0.0, F, fz, F, 1.0])
    rp, col, kap = gg.build_csr(6, e, np.zeros(len(e)))
    rows = [col[rp[x]:rp[x + 1]] for x in range(6)]
    a = orc.OracleGraph(rp, col, kap).run_one(0, 1, hop=6, flags=0, fval=f)
    up = np.flatnonzero(a["pkind"] == orc.K_UP)
    assert len(up) == 1 and a["pbirth"][up[0]] == F and a["pdeath"][up[0]] == fz     # the reference does emit it
    assert not certificate(6, rows, f, 0, 1)                                           # ... and the certificate declines
