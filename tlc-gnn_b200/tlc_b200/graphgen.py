"""Seeded synthetic graphs of the shapes BASELINE.json names.

All graphs are simple, undirected, without self-loops, Chung-Lu power-law with the named
node/edge counts, node ids shuffled.  Curvature kappa is per undirected edge; the benchmark
configs quantise it to a 1/1024 grid so that every path sum is exact in float64
-- `continuous=True` gives the un-quantised variant (parity case C2b).

Seeds: graph seed = cfg id, kappa seed = 100 + cfg id, negatives seed = 200 + cfg id.
"""
import numpy as np

SHAPES = {
    # name: (cfg id, N, M, gamma, max_degree_cap)
    "cora": (1, 2708, 5278, 2.9, None),
    "pubmed": (2, 19717, 44324, 2.8, None),
    "computers": (3, 13752, 245861, 2.3, 3000),
    "ppi": (4, 2400, 34000, 2.6, None),
    "collab": (5, 235868, 1285465, 2.6, None),
}


def chung_lu(N, M, gamma, seed, max_deg=None):
    """exactly M distinct undirected edges (lo<hi) as int64[M,2], expected degrees ~ power law."""
    rng = np.random.default_rng(seed)
    w = (np.arange(N, dtype=np.float64) + 1.0) ** (-1.0 / (gamma - 1.0))
    if max_deg is not None:
        # cap the expected degree 2M*w/sum(w): clipping renormalises, so iterate to a fixpoint
        for _ in range(50):
            w = np.minimum(w, max_deg * w.sum() / (2.0 * M))
    p = w / w.sum()
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    perm = rng.permutation(N)
    keys = np.empty(0, dtype=np.int64)
    while keys.size < M:
        need = int((M - keys.size) * 1.3) + 64
        a = np.searchsorted(cdf, rng.random(need), side="right")
        b = np.searchsorted(cdf, rng.random(need), side="right")
        ok = a != b
        a, b = perm[a[ok]], perm[b[ok]]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        new = lo.astype(np.int64) * N + hi
        # keep first occurrences, in draw order, not already present
        _, first = np.unique(new, return_index=True)
        new = new[np.sort(first)]
        new = new[~np.isin(new, keys)]
        keys = np.concatenate([keys, new])[:M] if keys.size + new.size >= M else np.concatenate([keys, new])
    keys = keys[:M]
    return np.stack([keys // N, keys % N], axis=1)


def make_kappa(M, seed, continuous=False, zero=False):
    """per-undirected-edge curvature.  zero -> hop-distance filtration (weight = 1)."""
    if zero:
        return np.zeros(M, dtype=np.float64)
    rng = np.random.default_rng(seed)
    k = rng.uniform(-0.9, 0.9, size=M)
    if not continuous:
        k = np.round(k * 1024.0) / 1024.0
    return k


def make_config(name, n_graphs=1, continuous=False, scale=1.0):
    """returns dict(edges int64[M,2], kappa f64[M], N).  `ppi` stacks n_graphs block-diagonally.
    `scale` shrinks N and M proportionally (parity tests on reduced shapes)."""
    cfg, N, M, gamma, cap = SHAPES[name]
    N = max(8, int(round(N * scale)))
    M = max(8, int(round(M * scale)))
    if cap is not None:
        cap = max(8, int(round(cap * max(scale, 0.05))))
    blocks, kap = [], []
    for gidx in range(n_graphs):
        e = chung_lu(N, M, gamma, seed=cfg + 1000 * gidx, max_deg=cap)
        blocks.append(e + gidx * N)
        kap.append(make_kappa(M, seed=100 + cfg + 1000 * gidx, continuous=continuous, zero=(name == "cora")))
    return dict(name=name, N=N * n_graphs, edges=np.concatenate(blocks), kappa=np.concatenate(kap))


def negative_pairs(N, edges, count, seed):
    """seeded uniform non-adjacent, non-identical pairs (collab config: equal number of negatives)."""
    rng = np.random.default_rng(seed)
    have = set((edges[:, 0] * N + edges[:, 1]).tolist())
    out = []
    while len(out) < count:
        a = rng.integers(0, N, size=count)
        b = rng.integers(0, N, size=count)
        for x, y in zip(a.tolist(), b.tolist()):
            if x == y:
                continue
            lo, hi = (x, y) if x < y else (y, x)
            if lo * N + hi in have:
                continue
            out.append((x, y))
            if len(out) == count:
                break
    return np.asarray(out, dtype=np.int64)


def relabel_first_appearance(edges):
    """the reference's node numbering: nx.Graph().add_edges_from(edges) then
    convert_node_labels_to_integers -> ids in first-appearance order (riccidist2dgm.py:217-220,
    loaddatas.py:88-92).  returns (labels_in_new_order int64[Nn], new_edges int64[M,2])."""
    flat = np.asarray(edges, dtype=np.int64).reshape(-1)
    uniq, first = np.unique(flat, return_index=True)
    order = np.argsort(first, kind="stable")
    labels = uniq[order]
    rank = np.empty(uniq.size, dtype=np.int64)
    rank[order] = np.arange(uniq.size)
    new = rank[np.searchsorted(uniq, flat)].reshape(-1, 2)
    return labels, new


def build_csr(N, edges, kappa):
    """symmetric CSR with ascending neighbour ids: rowptr int32[N+1], col int32[2M], kap f64[2M]."""
    e = np.asarray(edges, dtype=np.int64)
    src = np.concatenate([e[:, 0], e[:, 1]])
    dst = np.concatenate([e[:, 1], e[:, 0]])
    kk = np.concatenate([kappa, kappa]).astype(np.float64)
    order = np.lexsort((dst, src))
    src, dst, kk = src[order], dst[order], kk[order]
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(rowptr, src + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr.astype(np.int32), dst.astype(np.int32), np.ascontiguousarray(kk)
