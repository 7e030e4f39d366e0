#!/usr/bin/env python
"""Design-time analysis (CPU, uses the oracle as a data source; not part of the product):
phase counts of the phased Dijkstra (kernel 1b) and cycle / conflict statistics of the loop sweep (kernel 3b)
on vicinities of the named synthetic shape.

    python scripts/analyze_vicinity.py computers 2 6
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tlc-gnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402
from tlc_b200 import graphgen as gg  # noqa: E402


def phases_in_out(n, adj, w, root, use_out):
    """number of phases / settled per phase of the Crauser IN (+OUT) criterion"""
    INF = float("inf")
    dist = np.full(n, INF)
    state = np.zeros(n, np.int8)  # 0 far 1 tent 2 done
    minw = np.array([w[x].min() if len(w[x]) else INF for x in range(n)])
    dist[root] = 0
    state[root] = 1
    ph = 0
    sizes = []
    while True:
        tent = np.nonzero(state == 1)[0]
        if tent.size == 0:
            break
        dmin = dist[tent].min()
        take = dist[tent] <= dmin + minw[tent]
        if use_out:
            L = (dist[tent] + minw[tent]).min()
            take |= dist[tent] <= L
        S = tent[take]
        state[S] = 2
        for x in S:
            nd = dist[x] + w[x]
            y = adj[x]
            better = nd < dist[y]
            if better.any():
                np.minimum.at(dist, y[better], nd[better])
                st = state[y[better]]
                state[y[better]] = np.where(st == 0, 1, st)
        ph += 1
        sizes.append(S.size)
    return ph, sizes


def loops_stats(d, B=32):
    n, m = d["n"], d["m"]
    elo, ehi = d["elo"], d["ehi"]
    arank = np.empty(m, np.int64)
    arank[d["ord_asc"]] = np.arange(m)
    neg, pos = d["neg"], d["pos"]
    par = np.full(n, -1)
    pr = np.full(n, -1)
    root = elo[neg[0]]
    par[root] = root
    nb = [[] for _ in range(n)]
    for e in neg:
        nb[elo[e]].append((ehi[e], e))
        nb[ehi[e]].append((elo[e], e))
    q = [root]
    depth = np.zeros(n, np.int64)
    for x in q:
        for y, e in nb[x]:
            if par[y] < 0:
                par[y] = x
                pr[y] = arank[e]
                depth[y] = depth[x] + 1
                q.append(y)
    d0 = depth.copy()

    def cycle(p0, p1):
        s0 = []
        x = p0
        seen = {}
        while True:
            seen[x] = len(s0)
            s0.append(x)
            if par[x] == x:
                break
            x = par[x]
        y = p1
        s1 = []
        while y not in seen:
            s1.append(y)
            y = par[y]
        lca = y
        return s0[:seen[lca]], s1, lca, len(s0)

    clen, toroot = [], []
    # speculation: batches of B against the frozen tree; commit the longest prefix without a TRUE conflict
    k = 0
    npos = len(pos)
    batches = 0
    committed_hist = []
    while k < npos:
        hi = min(npos, k + B)
        spec = []
        for j in range(k, hi):
            pe = pos[j]
            a0, a1, lca, L0 = cycle(elo[pe], ehi[pe])
            ranks = set(pr[x] for x in a0) | set(pr[x] for x in a1)
            spec.append(ranks)
        removed = set()
        j = k
        while j < hi:
            if spec[j - k] & removed:
                break
            pe = pos[j]
            p0, p1 = elo[pe], ehi[pe]
            a0, a1, lca, L0 = cycle(p0, p1)
            clen.append(len(a0) + len(a1))
            toroot.append(L0)
            best, bc, in0 = -1, -1, 0
            for x in a0:
                if pr[x] > best:
                    best, bc, in0 = pr[x], x, 1
            for x in a1:
                if pr[x] > best:
                    best, bc, in0 = pr[x], x, 0
            removed.add(best)
            node, nodec, rc = (p0, p1, arank[pe]) if in0 else (p1, p0, arank[pe])
            while True:
                tp, tr = par[node], pr[node]
                par[node] = nodec
                pr[node] = rc
                if node == bc:
                    break
                nodec, rc, node = node, tr, tp
            j += 1
        committed_hist.append(j - k)
        k = j
        batches += 1
    return dict(npos=npos, depth_mean=float(d0.mean()), depth_max=int(d0.max()), cyc_mean=float(np.mean(clen)) if clen else 0,
                cyc_max=int(np.max(clen)) if clen else 0, toroot_mean=float(np.mean(toroot)) if toroot else 0,
                batches=batches, commit_mean=float(np.mean(committed_hist)) if committed_hist else 0)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "computers"
    hop = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cnt = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    what = sys.argv[4] if len(sys.argv) > 4 else "both"
    c = gg.make_config(name)
    labels, ne = gg.relabel_first_appearance(c["edges"])
    csr = gg.build_csr(len(labels), ne, c["kappa"])
    og = orc.OracleGraph(*csr)
    rng = np.random.default_rng(7)
    for ti in rng.choice(len(ne), cnt, replace=False):
        u, v = ne[ti]
        d = og.run_one(u, v, hop=hop, flags=orc.F_NORM | orc.F_EXTENDED)
        n, m = d["n"], d["m"]
        if d["status"] != 0:
            continue
        line = "target (%d,%d) n=%d m=%d" % (u, v, n, m)
        if what in ("both", "dijkstra"):
            adj = [[] for _ in range(n)]
            w = [[] for _ in range(n)]
            for a, b, ww in zip(d["elo"], d["ehi"], d["ew"]):
                adj[a].append(b); adj[b].append(a); w[a].append(ww); w[b].append(ww)
            adj = [np.array(x, np.int64) for x in adj]
            w = [np.array(x) for x in w]
            p_in, s_in = phases_in_out(n, adj, w, d["lu"], False)
            p_io, s_io = phases_in_out(n, adj, w, d["lu"], True)
            line += " | phases IN=%d IN+OUT=%d (first sizes IN %s)" % (p_in, p_io, s_in[:12])
        if what in ("both", "loops"):
            for B in (32, 256):
                st = loops_stats(d, B)
                line += " | B=%d: %s" % (B, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
        print(line, flush=True)


if __name__ == "__main__":
    main()


def spec_validity(d, B=32):
    """speculative batches of B positive edges against the frozen tree; an edge is VALID when no vertex of its frozen cycle had
    its parent pointer rewritten by an earlier edge of the batch (conservative check), else it is recomputed in order."""
    n, m = d["n"], d["m"]
    elo, ehi = d["elo"], d["ehi"]
    arank = np.empty(m, np.int64)
    arank[d["ord_asc"]] = np.arange(m)
    neg, pos = d["neg"], d["pos"]
    par = np.full(n, -1)
    pr = np.full(n, -1)
    root = elo[neg[0]]
    par[root] = root
    nb = [[] for _ in range(n)]
    for e in neg:
        nb[elo[e]].append((ehi[e], e))
        nb[ehi[e]].append((elo[e], e))
    q = [root]
    for x in q:
        for y, e in nb[x]:
            if par[y] < 0:
                par[y] = x
                pr[y] = arank[e]
                q.append(y)

    def walk(p0, p1):
        s0 = []
        x = p0
        seen = {}
        while True:
            seen[x] = len(s0)
            s0.append(x)
            if par[x] == x:
                break
            x = par[x]
        y = p1
        s1 = []
        while y not in seen:
            s1.append(y)
            y = par[y]
        return s0[:seen[y]], s1, y, len(s0)

    npos = len(pos)
    valid = true_ok = 0
    cyc = []
    root_len = []
    for k0 in range(0, npos, B):
        hi = min(npos, k0 + B)
        frozen = []
        for j in range(k0, hi):
            pe = pos[j]
            a0, a1, lca, L0 = walk(elo[pe], ehi[pe])
            frozen.append((set(a0) | set(a1) | {lca}, set(pr[x] for x in a0) | set(pr[x] for x in a1)))
            root_len.append(L0)
        touched = set()
        removed = set()
        for j in range(k0, hi):
            vs, rk = frozen[j - k0]
            if not (vs & touched):
                valid += 1
            if not (rk & removed):
                true_ok += 1
            pe = pos[j]
            p0, p1 = elo[pe], ehi[pe]
            a0, a1, lca, L0 = walk(p0, p1)
            cyc.append(len(a0) + len(a1))
            best, bc, in0 = -1, -1, 0
            for x in a0:
                if pr[x] > best:
                    best, bc, in0 = pr[x], x, 1
            for x in a1:
                if pr[x] > best:
                    best, bc, in0 = pr[x], x, 0
            removed.add(best)
            node, nodec, rc = (p0, p1, arank[pe]) if in0 else (p1, p0, arank[pe])
            while True:
                tp, tr = par[node], pr[node]
                par[node] = nodec
                pr[node] = rc
                touched.add(node)
                if node == bc:
                    break
                nodec, rc, node = node, tr, tp
    return dict(npos=npos, valid_frac=valid / max(npos, 1), true_ok_frac=true_ok / max(npos, 1), cyc_mean=float(np.mean(cyc)) if cyc else 0,
                cyc_p95=float(np.percentile(cyc, 95)) if cyc else 0, cyc_max=max(cyc) if cyc else 0, root_path_mean=float(np.mean(root_len)) if root_len else 0)
