// k3v_sweep.cu -- kernel 3v: the ascending union-find sweep in VERTEX order, one warp per vicinity.
//
// Replaces the ascending half of Union_find (accelerated_PD.py:27-68; KD copy :25-70) and produces the
// identical pair sequence [old[large], old[max_node]] (same values, same vertex ids, same order) as
// Kruskal over the reference's (perturbed key, canonical index) edge order -- without ever sorting the
// m edges.  Kernel 2v cut the (value, id) vertex order into blocks whose owned edges are totally
// ordered block against block; here a warp walks the blocks in order:
//
//   trivial block (the overwhelmingly common case): all vertices of the block carry the same value,
//     every one of them has a neighbour in an earlier block, and all those neighbours sit in ONE
//     component R.  Each vertex then joins R through its first edge as a singleton whose value equals
//     the edge's max value: the reference's `if old[large] < old[max_node]` (:65) is false, no pair is
//     emitted, whatever the order of the block's edges.  parent[x] = R, done -- one ballot.
//   general block: every owned edge (x, y) can only merge if it is the first, in (key, index) order,
//     among the block's edges joining the same pair of components-at-block-start.  The warp keeps that
//     first edge per (x, component) ("representative"), sorts the handful of representatives by
//     (key, lo, hi) and runs the reference's union rule on them in order (:53-67).
//
// A target whose block needs more than REP_CAP representatives is flagged in tfb[] and redone by the
// edge-sorted kernels 2 + 3 (same results, different cost).  The union-find parents are ranks in shared
// memory (16-bit when n < 65536).  Connectivity (riccidist2dgm.py:318) falls out as merges == n - 1.
#include "tlc_common.cuh"
#include "tlc_sort.cuh"

namespace tlc {
namespace {

constexpr int REP_CAP = 64;
constexpr int SWEEP_WARPS = 1;  // one warp per CTA: the shared-memory footprint (parents of ONE vicinity) sets the residency

struct RepBuf {
  unsigned long long key[REP_CAP];
  int32_t lo[REP_CAP], hi[REP_CAP];  // local ids, lo < hi (canonical orientation: edge[0] = lo)
  int32_t root[REP_CAP];             // component (rank of its root) of the earlier endpoint at block start
  int32_t perm[REP_CAP];
};

template <typename PT>
__device__ __forceinline__ int find_root(PT* p, int x) {  // path halving  accelerated_PD.py:53-58
  for (;;) {
    const int px = (int)p[x];
    if (px == x) return x;
    const int gp = (int)p[px];
    p[x] = (PT)gp;
    x = gp;
  }
}

// ---- the lower-adjacency stream: kernel 2v laid the owned edges out in sweep order, so the warp reads ladj
// strictly front to back.  A per-warp shared-memory ring is kept RING-CHUNK entries ahead with cp.async.
constexpr int RING = 512, CHUNK = 128;

struct AdjStream {
  const uint32_t* src;
  uint32_t* ring;
  int m, fetched, arrived;
  __device__ __forceinline__ void issue(int lane) {
#pragma unroll
    for (int k = 0; k < CHUNK / 32; k++) {
      const int idx = fetched + k * 32 + lane;
      if (idx < m) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + (idx & (RING - 1)));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + idx) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    fetched += CHUNK;
  }
  // make entries [j0, j1) readable (j1 - j0 <= 32); entries below j0 are dead.  Uniform over the warp.
  __device__ __forceinline__ void ensure(int j0, int j1, int lane) {
    if (j0 >= fetched + CHUNK) {  // the reader skipped ahead (blocks handled off-stream): drop the gap
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      fetched = j0 / CHUNK * CHUNK;
      arrived = fetched;
    }
    if (fetched < m && fetched + CHUNK - j0 <= RING) {
      __syncwarp();  // every lane is done reading the slots about to be overwritten
      while (fetched < m && fetched + CHUNK - j0 <= RING) issue(lane);
    }
    if (j1 > arrived) {
      const int need_end = (j1 + CHUNK - 1) / CHUNK * CHUNK;
      const int later = (fetched - need_end) / CHUNK;  // groups issued after the one needed: may stay in flight
      if (later >= 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
      else if (later == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else if (later == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      arrived = fetched - (later >= 3 ? 3 : later) * CHUNK;
    }
  }
  __device__ __forceinline__ uint32_t at(int j) const { return ring[j & (RING - 1)]; }
};

__device__ __forceinline__ bool rep_less(unsigned long long ka, int la, int ha, unsigned long long kb, int lb, int hb) {
  if (ka != kb) return ka < kb;
  if (la != lb) return la < lb;  // canonical edge index order == lexicographic (lo, hi)
  return ha < hb;
}

template <typename PT>
__global__ void __launch_bounds__(SWEEP_WARPS * 32) sweep_kernel(Params p, ChunkView c, int cap) {
  extern __shared__ unsigned char dyn_raw[];
  __shared__ RepBuf reps_all[SWEEP_WARPS];
  __shared__ uint32_t ring_all[SWEEP_WARPS][RING];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t = blockIdx.x * SWEEP_WARPS + wid;
  if (t >= c.T) return;
  const int n = c.tn[t];
  if (n == 0 || c.tstatus[t] > TLC_ST_TRIVIAL) return;
  const int64_t vo = c.voff[t], eo = c.eoff[t], po = c.poff(t);
  const int32_t* __restrict__ vord = c.vord + vo;
  const int32_t* __restrict__ bfirst = c.bfirst + vo + t;
  const int32_t* __restrict__ bend = c.bend + vo + t;
  const int32_t* __restrict__ loff = c.loff + vo + t;
  const uint32_t* __restrict__ ladj = c.ladj + eo;
  const double* __restrict__ fval = c.fval + vo;
  PT* parent = n <= cap ? reinterpret_cast<PT*>(dyn_raw) + (size_t)wid * cap : reinterpret_cast<PT*>(c.vs2 + vo);
  RepBuf& rb = reps_all[wid];
  const bool keep0 = (p.flags & TLC_F_KEEP_ZERO) != 0;
  const int nb = c.tnb[t];
  const unsigned FULL = 0xffffffffu;

  for (int r = lane; r < n; r += 32) parent[r] = (PT)r;
  __syncwarp();

  AdjStream adj{ladj, ring_all[wid], c.tm[t], 0, 0};
  int np = 0, nmerge = 0;
  int ncomp = 0;  // components among the vertices of the blocks processed so far
  bool bail = false;
  int s = 0, a0 = 0;
  // block table, 32 blocks per coalesced load, the next group already in flight
  int cur_first = 0, cur_end = 0, nxt_first = 0, nxt_end = 0;
  {
    const int i = lane;
    if (i < nb) { nxt_first = bfirst[i + 1]; nxt_end = bend[i]; }
  }
  for (int b = 0; b < nb && !bail; b++) {
    if ((b & 31) == 0) {
      cur_first = nxt_first; cur_end = nxt_end;
      const int i = b + 32 + lane;
      if (i < nb) { nxt_first = bfirst[i + 1]; nxt_end = bend[i]; }
    }
    const int e = __shfl_sync(FULL, cur_first, b & 31) & 0x7fffffff;
    const int we = __shfl_sync(FULL, cur_end, b & 31);
    const bool distinct = we < 0;
    const bool all_out = ((we >> 30) & 1) != 0;  // every vertex of the block has a neighbour in an earlier block
    const int a1 = we & 0x3fffffff;

    // ---------------- trivial block test ----------------
    bool done = false;
    if (!distinct && !keep0 && all_out) {
      int R = -1;
      bool ok = true;
      if (ncomp == 1) {
        // the processed prefix is ONE component: every earlier-block neighbour is in it, no entry needs reading
        R = find_root(parent, 0);
        a0 = a1;
      }
      // UNROLL chunks of 32 entries per step: the finds of a step are independent chains (two dependent
      // shared-memory loads each in the common case), issued together to overlap their latency
      constexpr int UNROLL = 4;
      for (int j0 = a0; j0 < a1 && ok; j0 += 32 * UNROLL) {
        uint32_t w[UNROLL];
        int p1[UNROLL], p2[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
          const int c0 = j0 + 32 * k;
          if (c0 < a1) adj.ensure(c0, min(c0 + 32, a1), lane);  // uniform
          const int j = c0 + lane;
          w[k] = j < a1 ? adj.at(j) : 0u;
        }
#pragma unroll
        for (int k = 0; k < UNROLL; k++) p1[k] = (w[k] >> 31) ? (int)parent[w[k] & 0x7fffffffu] : -1;
#pragma unroll
        for (int k = 0; k < UNROLL; k++) p2[k] = p1[k] >= 0 ? (int)parent[p1[k]] : -1;
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
          const bool out = (w[k] >> 31) != 0;  // entries inside the block are cycle edges once everything hangs off R
          int rt = p2[k];
          if (out && p2[k] != p1[k]) rt = find_root(parent, p2[k]);  // deeper than two levels: walk (and halve) the rest
          const unsigned bo = __ballot_sync(FULL, out);
          if (bo) {
            if (R < 0) R = __shfl_sync(FULL, rt, __ffs(bo) - 1);
            ok = ok && __all_sync(FULL, !out || rt == R);
          }
        }
      }
      if (ok) {
        __syncwarp();
        for (int x = s + lane; x < e; x += 32) parent[x] = (PT)R;
        __syncwarp();
        nmerge += e - s;
        done = true;  // ncomp unchanged: the block's vertices joined an existing component
      }
    }

    // ---------------- general block ----------------
    if (!done) {
      const int merges_before = nmerge;
      int nrep = 0;
      for (int x = s; x < e && !bail; x++) {
        const int xrep0 = nrep;
        const int lx = vord[x];
        const double fx = fval[lx];
        const int xa = loff[x], xb = loff[x + 1];
        for (int j0 = xa; j0 < xb && !bail; j0 += 32) {
          const int j = j0 + lane;
          const bool valid = j < xb;
          int y = 0, ly = 0, ry = -1, lo = 0, hi = 0;
          unsigned long long K = 0;
          if (valid) {
            y = (int)(ladj[j] & 0x7fffffffu);
            ly = vord[y];
            ry = find_root(parent, y);  // no union has happened in this block yet: component at block start
            K = f64_to_ordered(key_asc(fx, fval[ly]));
            lo = min(lx, ly); hi = max(lx, ly);
          }
          // first edge, in (key, lo, hi) order, per component within this chunk
          bool is_min = valid;
          for (int i = 0; i < 32; i++) {
            const int ri = __shfl_sync(FULL, ry, i);
            const unsigned long long Ki = __shfl_sync(FULL, K, i);
            const int loi = __shfl_sync(FULL, lo, i), hii = __shfl_sync(FULL, hi, i);
            if (is_min && i != lane && ri == ry && ri >= 0 && rep_less(Ki, loi, hii, K, lo, hi)) is_min = false;
          }
          // merge with the representatives x already has (earlier chunks of the same vertex)
          bool append = false;
          if (is_min) {
            int found = -1;
            for (int q = xrep0; q < nrep; q++) if (rb.root[q] == ry) found = q;
            if (found >= 0) {
              if (rep_less(K, lo, hi, rb.key[found], rb.lo[found], rb.hi[found])) { rb.key[found] = K; rb.lo[found] = lo; rb.hi[found] = hi; }
            } else append = true;
          }
          const unsigned ba = __ballot_sync(FULL, append);
          const int add = __popc(ba);
          if (nrep + add > REP_CAP) { bail = true; break; }
          if (append) {
            const int q = nrep + __popc(ba & lanemask_lt());
            rb.key[q] = K; rb.lo[q] = lo; rb.hi[q] = hi; rb.root[q] = ry;
          }
          nrep += add;
          __syncwarp();
        }
      }
      if (!bail) {
        // order the representatives: perm[rank] = q
        for (int q = lane; q < nrep; q += 32) {
          const unsigned long long K = rb.key[q];
          const int lo = rb.lo[q], hi = rb.hi[q];
          int rk = 0;
          for (int i = 0; i < nrep; i++) rk += rep_less(rb.key[i], rb.lo[i], rb.hi[i], K, lo, hi) ? 1 : 0;
          rb.perm[rk] = q;
        }
        __syncwarp();
        if (lane == 0) {
          const int32_t* vrank = c.vrank + vo;
          for (int i = 0; i < nrep; i++) {
            const int q = rb.perm[i];
            const int a = rb.lo[q], bb = rb.hi[q];  // edge = [a, b], a < b (local ids)
            const int A = find_root(parent, vrank[a]), B = find_root(parent, vrank[bb]);
            if (A == B) continue;
            const int la = vord[A], lb = vord[B];  // local ids of the two roots
            const double fA = fval[la], fB = fval[lb];
            const bool a_small = fA <= fB;         // small = pu if new[pu] <= new[pv]   :61-63
            const int small = a_small ? A : B, large = a_small ? B : A;
            const int llarge = a_small ? lb : la;
            const double flarge = a_small ? fB : fA;
            const double fa = fval[a], fb = fval[bb];
            const int max_node = fa > fb ? a : bb;  // :64
            const double fmaxn = fa > fb ? fa : fb;
            if (keep0 || flarge < fmaxn) {          // :65 (KD :68-69: always)
              c.pkind[po + np] = TLC_K_UP;
              c.pbv[po + np] = llarge; c.pdv[po + np] = max_node;
              c.pbirth[po + np] = flarge; c.pdeath[po + np] = fmaxn;
              np++;
            }
            parent[large] = (PT)small;            // :67
            nmerge++;
          }
        }
        np = __shfl_sync(FULL, np, 0);
        nmerge = __shfl_sync(FULL, nmerge, 0);
        __syncwarp();
        ncomp += (e - s) - (nmerge - merges_before);
      }
    }
    s = e;
    a0 = a1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  if (lane == 0) {
    if (bail) {
      c.tfb[t] = 1;
      if (c.fb_counter) atomicAdd(c.fb_counter, 1);
    } else {
      const int lmin = c.tminv[t], lmax = c.tmaxv[t];
      c.pkind[po + np] = TLC_K_ESS;               // [min_value, max_value]   accelerated_PD.py:110
      c.pbv[po + np] = lmin; c.pdv[po + np] = lmax;
      c.pbirth[po + np] = fval[lmin]; c.pdeath[po + np] = fval[lmax];
      c.tnp[t] = np + 1;
      c.tnneg[t] = 0; c.tnpos[t] = 0;
      if (nmerge != n - 1) c.tstatus[t] = TLC_ST_DISCONNECTED;  // assert len(components) == 1   riccidist2dgm.py:318
    }
  }
}

}  // namespace

void launch_sweep(const Params& p, const ChunkView& c, int64_t n_max, cudaStream_t st) {
  // parents in shared memory when SWEEP_WARPS vicinities of the chunk's largest size fit
  const bool narrow = n_max < 65536;
  const size_t esz = narrow ? 2 : 4;
  const size_t budget = 220 * 1024;
  int cap = (int)((n_max + 7) / 8 * 8);
  if ((size_t)cap * esz * SWEEP_WARPS > budget) cap = 0;
  const size_t bytes = (size_t)cap * esz * SWEEP_WARPS;
  const int grid = (c.T + SWEEP_WARPS - 1) / SWEEP_WARPS;
  if (narrow) {
    cudaFuncSetAttribute((const void*)sweep_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    sweep_kernel<uint16_t><<<grid, SWEEP_WARPS * 32, bytes, st>>>(p, c, cap);
  } else {
    cudaFuncSetAttribute((const void*)sweep_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    sweep_kernel<int32_t><<<grid, SWEEP_WARPS * 32, bytes, st>>>(p, c, cap);
  }
  count_launch();
}

}  // namespace tlc
