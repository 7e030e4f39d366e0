// k1b_filtration.cu -- kernel 1b: the node filtration of every vicinity.
//
// Replaces filtration.build_fv(weight_graph=True, norm) (riccidist2dgm.py:20-61) and the KD copy
// (Knowledge_Distillation/data_utils_NC.py:34-55).  The reference runs one nx.dijkstra_path per
// (vertex, root) and sums kappa+1 along the returned path with python's sum(); here each root gets ONE
// shortest-path computation inside the vicinity:
//   1. distances: edge-parallel Bellman-Ford relaxations with 64-bit atomicMin on the ordered bit
//      pattern of the (non-negative) float64 distance -- the least fixpoint of d[x] = min fl(d[y]+w),
//      i.e. exactly what Dijkstra computes; both roots relax in the same sweep over the edge list;
//   2. shortest-path tree: parent[x] = smallest local id y with fl(d[y] + w(y,x)) == d[x]
//      (packed (y, edge) 64-bit atomicMin) -- the same rule as the oracle's fast mode;
//   3. the path sum is re-accumulated from x towards the root in python-sum order (Neumaier
//      compensated as CPython >= 3.12 does, or plain with TLC_F_SUM_PLAIN)  -- SURVEY.md F5;
//   4. min / max / sum descriptors, block max-reduce, true division by the normaliser (:50-56).
// Precondition (as for networkx's Dijkstra): kappa + 1 > 0.
#include "tlc_common.cuh"

namespace tlc {
namespace {

constexpr unsigned long long INF_BITS = 0x7ff0000000000000ull;

struct PySum {  // CPython's float sum(): first item exact, then Neumaier (3.12+) or plain adds
  double s, c;
  int k;
};
__device__ __forceinline__ void pysum_add(PySum& p, double x, bool plain) {
  if (p.k == 0) { p.s = x; p.k = 1; return; }
  if (plain) { p.s = __dadd_rn(p.s, x); return; }
  const double t = __dadd_rn(p.s, x);
  if (fabs(p.s) >= fabs(x)) p.c = __dadd_rn(p.c, __dadd_rn(__dadd_rn(p.s, -t), x));
  else p.c = __dadd_rn(p.c, __dadd_rn(__dadd_rn(x, -t), p.s));
  p.s = t;
}
__device__ __forceinline__ double pysum_get(const PySum& p, bool plain) {
  if (p.k == 0) return 0.0;
  if (!plain && p.c != 0.0 && isfinite(p.c)) return __dadd_rn(p.s, p.c);
  return p.s;
}

__global__ void filtration_kernel(Params p, ChunkView c) {
  __shared__ double shd[32];
  __shared__ int sh_changed;
  const int t = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = c.tn[t], m = c.tm[t];
  if (n == 0) return;
  if (c.tstatus[t] > TLC_ST_TRIVIAL) return;
  const int64_t vo = c.voff[t], eo = c.eoff[t];
  const int32_t* __restrict__ elo = c.elo + eo;
  const int32_t* __restrict__ ehi = c.ehi + eo;
  const double* __restrict__ ew = c.ew + eo;
  double* d1 = c.d1 + vo;
  double* d2 = c.d2 + vo;
  double* fval = c.fval + vo;
  const int lu = c.tlu[t], lv = c.tlv[t];
  const bool node_mode = p.mode == TLC_MODE_NODE;
  const bool roots_in = lu >= 0 && lv >= 0;
  const bool plain = (p.flags & TLC_F_SUM_PLAIN) != 0;
  const bool two = roots_in && !node_mode && lu != lv;

  if (!roots_in) {
    // nx.NodeNotFound for every vertex -> dist = 100   riccidist2dgm.py:31-32,36-37
    for (int x = tid; x < n; x += nt) { d1[x] = 100.0; d2[x] = 100.0; }
  } else {
    unsigned long long* da = c.v64a + vo;  // ordered bits of dist to lu
    unsigned long long* db = c.v64b + vo;  // ... to lv
    for (int x = tid; x < n; x += nt) { da[x] = INF_BITS; db[x] = INF_BITS; }
    __syncthreads();
    if (tid == 0) { da[lu] = 0ull; db[lv] = 0ull; }
    __syncthreads();
    // 1. relaxations until a sweep changes nothing (<= n sweeps: guards kappa+1 <= 0 misuse)
    for (int round = 0; round < n + 1; round++) {
      if (tid == 0) sh_changed = 0;
      __syncthreads();
      int ch = 0;
      for (int e = tid; e < m; e += nt) {
        const int a = elo[e], b = ehi[e];
        const double w = ew[e];
        {
          const double xa = __longlong_as_double((long long)da[a]), xb = __longlong_as_double((long long)da[b]);
          const double ta = __dadd_rn(xa, w), tb = __dadd_rn(xb, w);
          if (ta < xb) { atomicMin(&da[b], (unsigned long long)__double_as_longlong(ta)); ch = 1; }
          else if (tb < xa) { atomicMin(&da[a], (unsigned long long)__double_as_longlong(tb)); ch = 1; }
        }
        if (two) {
          const double xa = __longlong_as_double((long long)db[a]), xb = __longlong_as_double((long long)db[b]);
          const double ta = __dadd_rn(xa, w), tb = __dadd_rn(xb, w);
          if (ta < xb) { atomicMin(&db[b], (unsigned long long)__double_as_longlong(ta)); ch = 1; }
          else if (tb < xa) { atomicMin(&db[a], (unsigned long long)__double_as_longlong(tb)); ch = 1; }
        }
      }
      if (ch) sh_changed = 1;
      __syncthreads();
      const int any = sh_changed;
      __syncthreads();
      if (!any) break;
    }
    // 2.+3. per root: tree by packed (parent, edge) atomicMin, then python-order path sums
    unsigned long long* key = c.v64c + vo;  // packed (parent, edge) of the shortest-path tree
    for (int r = 0; r < (two ? 2 : 1); r++) {
      unsigned long long* dist = r == 0 ? da : db;
      const int root = r == 0 ? lu : lv;
      double* out = r == 0 ? d1 : d2;
      for (int x = tid; x < n; x += nt) key[x] = ~0ull;
      __syncthreads();
      for (int e = tid; e < m; e += nt) {
        const int a = elo[e], b = ehi[e];
        const double w = ew[e];
        const double xa = __longlong_as_double((long long)dist[a]), xb = __longlong_as_double((long long)dist[b]);
        if (__dadd_rn(xa, w) == xb) atomicMin(&key[b], ((unsigned long long)(uint32_t)a << 32) | (uint32_t)e);
        if (__dadd_rn(xb, w) == xa) atomicMin(&key[a], ((unsigned long long)(uint32_t)b << 32) | (uint32_t)e);
      }
      __syncthreads();
      for (int x = tid; x < n; x += nt) {
        double res;
        if (x == root) res = 0.0;
        else if (dist[x] == INF_BITS) res = 100.0;  // nx.NetworkXNoPath -> 100 (disconnected vicinity; status 3 later)
        else {
          PySum ps{0.0, 0.0, 0};
          int y = x, guard = 0;
          while (y != root && guard++ <= n) {
            const unsigned long long k = key[y];
            pysum_add(ps, ew[(uint32_t)k], plain);  // ricci_curv[(path[y], path[y+1])] + 1   :30
            y = (int)(k >> 32);
          }
          res = pysum_get(ps, plain);
        }
        out[x] = res;
      }
      __syncthreads();
    }
    if (!two) for (int x = tid; x < n; x += nt) d2[x] = d1[x];
    __syncthreads();
    // `if x in [root_1, root_2]`: all three attributes 0   riccidist2dgm.py:22-25
    if (tid == 0) { d1[lu] = 0.0; d2[lu] = 0.0; d1[lv] = 0.0; d2[lv] = 0.0; }
  }
  __syncthreads();

  // 4. descriptors + normalisation   riccidist2dgm.py:47-56 ; data_utils_NC.py:52-54
  double mx = -1.0, sm = -1.0;
  for (int x = tid; x < n; x += nt) {
    const double a = d1[x], b = d2[x];
    mx = fmax(mx, fmax(a, b));
    sm = fmax(sm, node_mode ? a : __dadd_rn(a, b));
  }
  double smax = block_reduce_max(mx, shd);
  double ssum = block_reduce_max(sm, shd);
  const bool norm = (p.flags & TLC_F_NORM) != 0;
  if (norm) {
    if (p.flags & TLC_F_NORM_EPS) { smax = __dadd_rn(smax, 1e-10); ssum = __dadd_rn(ssum, 1e-10); }
    else if (smax == 0.0 || ssum == 0.0) {  // ZeroDivisionError -> zeros   riccidist2dgm.py:54-56,356-357
      if (tid == 0) c.tstatus[t] = TLC_ST_DEGENERATE;
      return;
    }
  }
  for (int x = tid; x < n; x += nt) {
    const double a = d1[x], b = d2[x];
    double f;
    if (p.descriptor == TLC_DESC_MIN) f = fmin(a, b);
    else if (p.descriptor == TLC_DESC_MAX) f = fmax(a, b);
    else f = node_mode ? a : __dadd_rn(a, b);
    if (norm) f = __ddiv_rn(f, p.descriptor == TLC_DESC_SUM ? ssum : smax);
    fval[x] = f;
  }
}

}  // namespace

void launch_filtration(const Params& p, const ChunkView& c, int block, cudaStream_t st) {
  filtration_kernel<<<c.T, block, 0, st>>>(p, c);
  count_launch();
}

}  // namespace tlc
