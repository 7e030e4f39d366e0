// k3b_loops.cu -- kernel 3b: 1-dim extended persistence pairs by loop tracing, one CTA per vicinity.
//
// Replaces Accelerate_PD (accelerated_PD.py:115-178; KD copy Knowledge_Distillation/accelerated_PD.py:120-183).
// The spanning tree of the descending sweep's Neg edges is rooted at the first endpoint of the first
// Neg edge (:119-125).  Positive edges are then processed strictly in sweep order: the tree cycle they
// close is path(p0 -> lca) + path(p1 -> lca) (== the reference's symmetric difference of the two root
// paths, :141-151); the cycle edge with the largest `asc` key is paired with the positive edge
// (:155-165) and swapped out of the tree by reversing parent pointers from the positive edge's
// endpoint up to that edge's child end (:168-176).
//
// `asc` comparisons use the edge's RANK in the ascending sweep (kernel 2), a total order that agrees
// with the float64 key; among equal keys the reference takes the first maximum in python-set
// iteration order, which is implementation-defined and does not change any pair VALUE --
// the canonical choice here and in the oracle is the latest edge of the ascending sweep.
//
// Work split: the whole CTA computes edge ranks, roots the tree (parallel relaxation over the n-1
// tree edges) and, afterwards, turns the recorded (loop-max edge, positive edge) couples into pairs
// with an ordered compaction.  The sweep itself is sequential by nature and runs on warp 0 with the
// per-vertex state (parent, rank of the parent edge, visit stamp) in shared memory; the next 32
// positive edges are gathered by the 32 lanes at once so the sequential part never waits on HBM.
//
// Two schedules of the sequential sweep.  Large vicinities: one CTA per vicinity, state in shared memory, lane 0 walks.
// Small vicinities (the reference's own settings: hop 1, or sparse graphs): ONE LANE per vicinity
// (loops_sweep_lanes_kernel), each lane's state in its own slice of shared memory -- up to several hundred independent
// sequential sweeps in flight per SM instead of one per resident CTA (at most 32); the parallel phases before and
// after the sweep still run a CTA per vicinity (phase mask).
#include "tlc_common.cuh"

namespace tlc {
namespace {

constexpr int PH_PREP = 1, PH_SWEEP = 2, PH_EMIT = 4;

// one positive edge (p0, p1) of rank `rc`, the k-th of the sweep: returns the child end of the loop-max edge (its parent
// end through *by) and re-hangs the tree.  tpar / tpr / stamp: parent, rank of the parent edge, visit stamp per vertex.
template <typename PT>
__device__ __forceinline__ int sweep_one_edge(PT* tpar, int32_t* tpr, int32_t* stamp, int k, int p0, int p1, int rc, int* by) {
  // path_0: p0 -> root, stamped                                                      accelerated_PD.py:131-144
  for (int x = p0;;) { stamp[x] = k; const int px = (int)tpar[x]; if (px == x) break; x = px; }
  int lca = p1;                                                                       // :145-151
  while (stamp[lca] != k) lca = (int)tpar[lca];
  int best = -1, bc = -1, in0 = 0;
  for (int x = p0; x != lca; x = (int)tpar[x]) { const int r = tpr[x]; if (r > best) { best = r; bc = x; in0 = 1; } }
  for (int x = p1; x != lca; x = (int)tpar[x]) { const int r = tpr[x]; if (r > best) { best = r; bc = x; in0 = 0; } }
  *by = (int)tpar[bc];                                                                // large_edge   :155-159
  // change the parent                                                                :168-176
  int node = in0 ? p0 : p1, nodec = in0 ? p1 : p0;
  for (;;) {
    const int tp = (int)tpar[node], tr = tpr[node];
    tpar[node] = (PT)nodec; tpr[node] = rc;
    if (node == bc) break;
    nodec = node; rc = tr; node = tp;
  }
  return bc;
}

struct LoopShared {
  int32_t wcnt[33];
  int32_t changed;
  int32_t g0[32], g1[32], gr[32];
};

__global__ void loops_kernel(Params p, ChunkView c, int smem_ints, int phases) {
  extern __shared__ int32_t dyn[];
  __shared__ LoopShared sh;
  const int t = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const int n = c.tn[t], m = c.tm[t];
  if (n == 0 || c.tstatus[t] > TLC_ST_TRIVIAL) return;
  const int npos = c.tnpos[t], nneg = c.tnneg[t];
  if (nneg == 0) {  // list(Nodes)[0] -> IndexError   accelerated_PD.py:122  (single-vertex vicinity)
    if (tid == 0) c.tstatus[t] = TLC_ST_NO_TREE_EDGES;
    return;
  }
  const int64_t vo = c.voff[t], eo = c.eoff[t], po = c.poff(t);
  const int32_t* __restrict__ elo = c.elo + eo;
  const int32_t* __restrict__ ehi = c.ehi + eo;
  const uint32_t* __restrict__ ord_asc = c.ord_asc + eo;
  const int32_t* __restrict__ pos = c.pos + eo;
  const int32_t* __restrict__ neg = c.neg + vo;
  int32_t* arank = c.arank + eo;
  const bool in_smem = 3 * n <= smem_ints;
  int32_t* tpar = in_smem ? dyn : c.vs0 + vo;
  int32_t* tpr = in_smem ? dyn + n : c.vs1 + vo;
  int32_t* stamp = in_smem ? dyn + 2 * n : c.vs2 + vo;
  uint32_t* out_x = c.sp0 + eo;  // per positive edge: child end of the loop-max edge
  uint32_t* out_y = c.sp1 + eo;  // ... and its parent end

  if (phases & PH_PREP) {
  for (int k = tid; k < m; k += nt) arank[ord_asc[k]] = k;
  for (int x = tid; x < n; x += nt) { tpar[x] = -1; stamp[x] = -1; }
  __syncthreads();
  // root the tree: Parent from nx.bfs_tree(g, root)   accelerated_PD.py:119-125
  if (tid == 0) { const int root = elo[neg[0]]; tpar[root] = root; tpr[root] = -1; }
  __syncthreads();
  for (int round = 0; round < n; round++) {
    if (tid == 0) sh.changed = 0;
    __syncthreads();
    int ch = 0;
    for (int i = tid; i < nneg; i += nt) {
      const int e = neg[i];
      const int a = elo[e], b = ehi[e];
      const int pa = tpar[a], pb = tpar[b];
      if (pa >= 0 && pb < 0) { tpar[b] = a; tpr[b] = arank[e]; ch = 1; }
      else if (pb >= 0 && pa < 0) { tpar[a] = b; tpr[a] = arank[e]; ch = 1; }
    }
    if (ch) sh.changed = 1;
    __syncthreads();
    const int any = sh.changed;
    __syncthreads();
    if (!any) break;
  }
  }  // PH_PREP

  // sequential sweep over the positive edges: the lanes of warp 0 gather the next 32 positive edges
  // (endpoints, rank) into shared memory, lane 0 walks the tree
  if ((phases & PH_SWEEP) && wid == 0) {
    for (int k0 = 0; k0 < npos; k0 += 32) {
      if (k0 + lane < npos) {
        const int pe = pos[k0 + lane];
        sh.g0[lane] = elo[pe]; sh.g1[lane] = ehi[pe]; sh.gr[lane] = arank[pe];
      }
      __syncwarp();
      if (lane == 0) {
        const int cnt = min(32, npos - k0);
        for (int j = 0; j < cnt; j++) {
          const int k = k0 + j;
          int y = 0;
          const int x = sweep_one_edge<int32_t>(tpar, tpr, stamp, k, sh.g0[j], sh.g1[j], sh.gr[j], &y);
          out_x[k] = (uint32_t)x; out_y[k] = (uint32_t)y;
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (!(phases & PH_EMIT)) return;

  // pairs [low_value, large_value] in positive-edge order, ordered compaction          :160-165
  const bool keep0 = (p.flags & TLC_F_KEEP_ZERO) != 0;
  const int32_t* __restrict__ vcls = c.vcls + vo;
  const double* __restrict__ fval = c.fval + vo;
  int cursor = c.tnp[t];
  for (int base = 0; base < npos; base += nt) {
    const int k = base + tid;
    bool emit = false;
    int lo_v = 0, lv_v = 0;
    if (k < npos) {
      const int pe = pos[k];
      const int p0 = elo[pe], p1 = ehi[pe];
      const int x = (int)out_x[k], y = (int)out_y[k];
      const int la = min(x, y), lb = max(x, y);
      lv_v = vcls[la] >= vcls[lb] ? la : lb;   // large_value = max(old[large_edge])
      lo_v = vcls[p0] <= vcls[p1] ? p0 : p1;   // low_value  = min(old[pos_edge])
      emit = keep0 || vcls[lv_v] > vcls[lo_v];
    }
    const unsigned bal = __ballot_sync(0xffffffffu, emit);
    __syncthreads();
    if (lane == 0) sh.wcnt[wid] = __popc(bal);
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < nw; w++) { const int cw = sh.wcnt[w]; if (w < wid) pre += cw; tot += cw; }
    if (emit) {
      const int64_t o = po + cursor + pre + __popc(bal & lanemask_lt());
      c.pkind[o] = TLC_K_ONE;
      c.pbv[o] = lo_v; c.pdv[o] = lv_v;
      c.pbirth[o] = fval[lo_v]; c.pdeath[o] = fval[lv_v];
    }
    cursor += tot;
  }
  if (tid == 0) c.tnp[t] = cursor;
}

// the sequential sweep of SMALL vicinities: one lane per vicinity, state in the lane's slice of shared memory
// (parent u16, rank of the parent edge, stamp: 10 bytes per vertex, `cap` vertices per lane)
__global__ void __launch_bounds__(128) loops_sweep_lanes_kernel(ChunkView c, int cap) {
  extern __shared__ int32_t dyn[];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c.T) return;
  const int n = c.tn[t];
  if (n == 0 || c.tstatus[t] > TLC_ST_TRIVIAL) return;
  const int npos = c.tnpos[t];
  if (c.tnneg[t] == 0) return;  // (status 7 set by the preparation phase)
  const int64_t vo = c.voff[t], eo = c.eoff[t];
  const int32_t* __restrict__ elo = c.elo + eo;
  const int32_t* __restrict__ ehi = c.ehi + eo;
  const int32_t* __restrict__ pos = c.pos + eo;
  const int32_t* __restrict__ arank = c.arank + eo;
  int32_t* tpr = dyn + (size_t)threadIdx.x * cap;
  int32_t* stamp = dyn + (size_t)blockDim.x * cap + (size_t)threadIdx.x * cap;
  uint16_t* tpar = reinterpret_cast<uint16_t*>(dyn + (size_t)2 * blockDim.x * cap) + (size_t)threadIdx.x * cap;
  {  // the rooted tree of the preparation phase
    const int32_t* gpar = c.vs0 + vo;
    const int32_t* gpr = c.vs1 + vo;
    for (int x = 0; x < n; x++) { tpar[x] = (uint16_t)gpar[x]; tpr[x] = gpr[x]; stamp[x] = -1; }
  }
  uint32_t* out_x = c.sp0 + eo;
  uint32_t* out_y = c.sp1 + eo;
  // the next positive edge is fetched while the current one walks the tree
  int pe = npos > 0 ? pos[0] : 0;
  int p0 = npos > 0 ? elo[pe] : 0, p1 = npos > 0 ? ehi[pe] : 0, rc = npos > 0 ? arank[pe] : 0;
  for (int k = 0; k < npos; k++) {
    int q0 = 0, q1 = 0, qr = 0;
    if (k + 1 < npos) { const int pn = pos[k + 1]; q0 = elo[pn]; q1 = ehi[pn]; qr = arank[pn]; }
    int y = 0;
    const int x = sweep_one_edge<uint16_t>(tpar, tpr, stamp, k, p0, p1, rc, &y);
    out_x[k] = (uint32_t)x; out_y[k] = (uint32_t)y;
    p0 = q0; p1 = q1; rc = qr;
  }
}

}  // namespace

void launch_loops(const Params& p, const ChunkView& c, int block, int smem_ints, int64_t n_max, cudaStream_t st) {
  // small vicinities (and enough of them to matter): CTA-parallel preparation, lane-per-vicinity sweep, CTA-parallel emit
  const int cap = (int)((n_max + 1) / 2 * 2);                        // (even: keeps the u16 slices 4-byte aligned)
  int lanes = cap > 0 ? (int)((size_t)200 * 1024 / ((size_t)cap * 10)) / 32 * 32 : 0;
  if (lanes > 128) lanes = 128;
  if (lanes >= 64 && c.T >= 2048) {
    const size_t bytes = (size_t)lanes * cap * 10;
    loops_kernel<<<c.T, block, 0, st>>>(p, c, 0, PH_PREP);
    count_launch();
    cudaFuncSetAttribute((const void*)loops_sweep_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    loops_sweep_lanes_kernel<<<(c.T + lanes - 1) / lanes, lanes, bytes, st>>>(c, cap);
    count_launch();
    loops_kernel<<<c.T, block, 0, st>>>(p, c, 0, PH_EMIT);
    count_launch();
    return;
  }
  const size_t bytes = (size_t)smem_ints * 4;
  if (bytes > 48 * 1024)
    cudaFuncSetAttribute((const void*)loops_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  loops_kernel<<<c.T, block, bytes, st>>>(p, c, smem_ints, PH_PREP | PH_SWEEP | PH_EMIT);
  count_launch();
}

}  // namespace tlc
