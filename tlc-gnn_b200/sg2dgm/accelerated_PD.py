"""Mirror of sg2dgm/accelerated_PD.py (and, with kd=True, Knowledge_Distillation/accelerated_PD.py).

Same three functions, same data shapes:
  perturb_filter_function(g, descriptor)            accelerated_PD.py:6-23   -> simplex_filter dict
  Union_find(simplex_filter)                        accelerated_PD.py:26-113 -> (PD, Pos_edges, Neg_edges)
  Accelerate_PD(Pos_edges, Neg_edges, simplex_filter) accelerated_PD.py:115-178 -> PD_one
The dict order of simplex_filter is the tie-break, exactly as python's stable sort makes it in the
reference.  Sorting, both union-find sweeps and the loop tracing run in kernels 2 / 3 / 3b.
"""
import numpy as np

from tlc_b200 import _lib as L
from tlc_b200 import api

ee = 1e-6
max_filter = 101


def perturb_filter_function(g, descriptor='seal'):
    """accelerated_PD.py:6-23.  `descriptor` may also be a list/array of filtration values indexed by
    node label (the KD signature, Knowledge_Distillation/accelerated_PD.py:6-13)."""
    simplex_filter = {}
    by_list = not isinstance(descriptor, str)
    for node in g.nodes():
        val = descriptor[node] if by_list else g.nodes[node][descriptor]
        simplex_filter[node] = {'old': val, 'new': val}
    for edge in g.edges():
        a, b = simplex_filter[edge[0]]['old'], simplex_filter[edge[1]]['old']
        max_node, min_node = max(a, b), min(a, b)
        simplex_filter[(edge[0], edge[1])] = {'asc': max_node + (min_node + 1) * ee,
                                              'desc': min_node - (max_filter - max_node) * ee}
    return simplex_filter


def _run(simplex_filter, extended, kd, device=0):
    nodes = [s for s in simplex_filter if not isinstance(s, tuple)]
    edges = [s for s in simplex_filter if isinstance(s, tuple)]
    idx = {x: i for i, x in enumerate(nodes)}
    fval = np.array([simplex_filter[x]['old'] for x in nodes], dtype=np.float64)
    e = np.array([(idx[a], idx[b]) for a, b in edges], dtype=np.int32).reshape(-1, 2)
    flags = (L.F_EXTENDED if extended else 0) | (L.F_KEEP_ZERO if kd else 0)
    r = api.union_find(fval, e, flags=flags, device=device)
    return nodes, edges, r


def Union_find(simplex_filter, kd=False, device=0):
    """accelerated_PD.py:26-113 -> (PD, Pos_edges, Neg_edges); with kd=True the KD return
    (PD_up, [[min,max]], PD_down, Pos_edges, Neg_edges) of Knowledge_Distillation/accelerated_PD.py:118."""
    nodes, edges, r = _run(simplex_filter, False, kd, device)
    pairs = np.stack([r["pbirth"], r["pdeath"]], 1)
    pos = [[edges[i][0], edges[i][1]] for i in r["pos"]]
    neg = [[edges[i][0], edges[i][1]] for i in r["neg"]]
    if kd:
        k = r["pkind"]
        return (pairs[k == L.K_UP], pairs[k == L.K_ESS], pairs[k == L.K_DOWN], pos, neg)
    return pairs.tolist(), pos, neg


def Accelerate_PD(Pos_edges, Neg_edges, simplex_filter, kd=False, device=0):
    """accelerated_PD.py:115-178 -> PD_one.  Pos/Neg are re-derived on the device from simplex_filter
    (they are a pure function of it); an empty Neg list raises IndexError as the reference does (:122)."""
    if len(Neg_edges) == 0:
        raise IndexError("list index out of range")
    nodes, edges, r = _run(simplex_filter, True, kd, device)
    one = r["pkind"] == L.K_ONE
    out = np.stack([r["pbirth"][one], r["pdeath"][one]], 1)
    return out if kd else out.tolist()
