// k2v_vorder.cu -- kernel 2v: the vertex order of every vicinity, cut into blocks.
//
// Replaces, together with kernel 3v, the ascending half of perturb_filter_function + Union_find
// (accelerated_PD.py:6-23, :27-68) WITHOUT sorting the m edges.  The reference sorts all simplices by
// the perturbed float64 key  asc(e) = fl(M + fl(fl(mu + 1) * 1e-6)),  M/mu = larger/smaller endpoint value.
// Every edge is OWNED by its later endpoint in the (value, local id) vertex order, and fl(+), fl(*) are
// monotone, so the keys of the edges owned by a vertex x lie in
//     [ lo(x) = fl(f_x + fl(fl(f_min + 1) * 1e-6)),  hi(x) = fl(f_x + fl(fl(f_x + 1) * 1e-6)) ].
// Cutting the sorted vertex sequence wherever hi(v_i) < lo(v_{i+1}) therefore yields BLOCKS such that
// every edge owned by an earlier block precedes, in the reference's (key, canonical index) order, every
// edge owned by a later block -- exactly, in IEEE arithmetic, including the perturbed-key crossings of
// SURVEY.md F4 (vertices closer than ~1e-6 simply share a block).  Kernel 3v then runs Kruskal block by
// block; only inside a block do individual edge keys matter.
//
// One CTA per vicinity:
//   0. the no-pair certificate (see step 0 in the kernel): where it holds the diagram is the essential pair alone, which is
//      written here, and steps 1-3 as well as kernel 3v are skipped for the target
//   1. stable LSD radix sort of the n vertices on the ordered image of their value rounded down to float,
//      runs of equal floats fixed to exact (float64 value, id) order (ties keep ascending local id)                    -> vord[rank] = local id, vrank[local id] = rank
//   2. block starts bfirst[b], with two flags kernel 3v decides on:
//        bit 31: the block holds distinct values (a near-tie block: edge keys inside it interleave)
//        bit 30: every vertex of the block has a neighbour in an EARLIER block (one early-exit scan of
//                its adjacency row against the packed rank/block table in shared memory)
//   3. local ids of the essential pair [min, max]        (accelerated_PD.py:35-38,110: first vertex in
//      ascending id attaining the extreme value)
#include <cstdlib>

#include "tlc_common.cuh"
#include "tlc_sort.cuh"

namespace tlc {
namespace {

__global__ void __launch_bounds__(512, 3) vorder_kernel(Params p, ChunkView c, int t0, int smem_ints, int bm_in_smem, int sort_cap) {
  extern __shared__ int32_t dyn[];
  __shared__ SortShared sh;
  const int t = t0 + blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = c.tn[t];
  if (tid == 0) c.tfb[t] = 0;
  if (n == 0 || c.tstatus[t] > TLC_ST_TRIVIAL) return;
  const int64_t vo = c.voff[t], ao = c.dbm ? 0 : c.aoff[t];
  const double* __restrict__ fval = c.fval + vo;
  int32_t* vord = c.vord + vo;
  int32_t* vrank = c.vrank + vo;
  int32_t* bfirst = c.bfirst + vo + t;

  // ---- 0. certificate: a filtration whose only local minimum is the root plateau has NO ordinary pair ----
  // With delta = 1e-8 call a vertex x (other than the roots)
  //   A  if it has a neighbour y with f_y < f_x - delta;
  //   B  if it is not A, has NO neighbour with f_x - delta <= f_y < f_x or f_x < f_y <= f_x + delta, and an exactly equal
  //      neighbour; B is GROUNDED if a chain of exactly equal neighbours leads to an A vertex.
  // If all values lie in [0, 1], the roots carry 0, are adjacent (or one), and every other vertex is A or grounded B, the
  // reference's sweep
  // (accelerated_PD.py:40-68) emits no pair at all:
  //   * the edge u-v has the smallest possible key 1e-6 and comes first;
  //   * the first edge of an A vertex x, in (key, position) order, is a descent edge to its lowest neighbour y* -- key
  //     fl(f_x + fl(fl(f_y* + 1) * 1e-6)) lies >= 45 ulps below the key of every edge in which x is the smaller or an
  //     equal end -- and by induction over the keys y* already hangs on the roots' component: x joins as a singleton of
  //     value f_x at an edge of larger value f_x, `old[large] < old[max_node]` (:65) is false;
  //   * the edges among equal values F all carry ONE key, above the first edges of the A vertices of value F and >= 1e-14
  //     below every edge from a B vertex of value F to a higher neighbour (delta apart); they merge components whose
  //     minima are F or 0 at edges of larger value F: false again; after them every grounded B vertex is attached.
  // The diagram is then the essential pair [min, max] alone and the vicinity is connected.  Distance-to-roots filtrations
  // look like this (the headline shape: one point per diagram), so the sort, the block cuts and the sweep are skipped for
  // such targets: kernel 3v sees tnb = -1.  Near-ties (the perturbed-key crossings of SURVEY.md F4) fail the tests above
  // and take the full path.  Only on the graph-row route (batch calls; diagram / detail calls run everything), never with
  // keep-zero pairs (KD flags).
  if (c.dbm && !(p.flags & TLC_F_KEEP_ZERO) && !c.no_fast) {
    __shared__ unsigned long long s_fmax;
    __shared__ int s_arg, s_adj, s_ncand;
    constexpr double DELTA = 1e-8;
    const int lu = c.tlu[t], lv = c.tlv[t];
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    int bad = (lu < 0 || lv < 0) ? 1 : 0;
    if (tid == 0) { s_fmax = 0ull; s_arg = 0x7fffffff; s_adj = (lu == lv) ? 1 : 0; s_ncand = 0; }
    __syncthreads();
    const int32_t* __restrict__ astart0 = c.astart + vo;
    const int32_t* __restrict__ adeg0 = c.adeg + vo;
    const int32_t* __restrict__ h1 = c.neg + vo;
    const int32_t* __restrict__ h0 = c.vcls + vo;
    const uint32_t* __restrict__ gbm0 = c.dbm + (size_t)t * 2 * c.W;
    const int32_t* __restrict__ gcol0 = c.gcol;
    int32_t* cand = c.vs1 + vo;   // vertices the tree-parent hints do not settle
    int32_t* vcl = c.vs2 + vo;    // 1: A or grounded B, 2: B not grounded yet
    if (!bad) {
      unsigned long long mymax = 0ull;
      for (int x0 = tid; x0 < n; x0 += 4 * nt) {
        double f4[4];
        int a4[4], b4[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int x = x0 + u * nt;
          f4[u] = x < n ? fval[x] : 0.0; a4[u] = x < n ? h1[x] : -1; b4[u] = x < n ? h0[x] : -1;
        }
        double fa4[4], fb4[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { fa4[u] = a4[u] >= 0 ? fval[a4[u]] : 2.0; fb4[u] = b4[u] >= 0 ? fval[b4[u]] : 2.0; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int x = x0 + u * nt;
          if (x >= n) continue;
          const double fx = f4[u];
          const unsigned long long ko = f64_to_ordered(fx);
          mymax = ko > mymax ? ko : mymax;
          if (x == lu || x == lv) { if (fx != 0.0) bad = 1; vcl[x] = 1; continue; }
          if (!(fx > 0.0)) { bad = 1; continue; }
          const double lim = fx - DELTA;
          if (fa4[u] < lim || fb4[u] < lim) vcl[x] = 1;
          else { vcl[x] = 0; cand[atomicAdd(&s_ncand, 1)] = x; }
        }
      }
      atomicMax(&s_fmax, mymax);
      if (lu != lv) {  // the two roots must be adjacent
        const int32_t gv = c.vert[vo + lv];
        const int a = astart0[lu], dg = adeg0[lu];
        int f = 0;
        for (int j = tid; j < dg; j += nt) f |= gcol0[a + j] == gv;
        if (f) s_adj = 1;
      }
    }
    bad = __syncthreads_or(bad);
    if (!bad && !s_adj) bad = 1;
    if (!bad && s_fmax > f64_to_ordered(1.0)) bad = 1;  // (the ulp margins of the proof are for values in [0, 1]: normalised filtrations)
    if (!bad) {
      // the candidates, a warp per row: A after all, B, or neither
      const int ncand = s_ncand;
      for (int i = wid; i < ncand; i += nw) {
        const int x = cand[i];
        const double fx = fval[x];
        const int a = astart0[x], dg = adeg0[x];
        int fl = 0;  // bit 0: lower by more than delta, 1: lower within delta, 2: equal, 3: higher within delta
        for (int j = lane; j < dg; j += 32) {
          const int y = bitmap_rank(gbm0, c.W, gcol0[a + j]);
          if (y < 0) continue;
          const double fy = fval[y];
          if (fy < fx - DELTA) fl |= 1;
          else if (fy < fx) fl |= 2;
          else if (fy == fx) fl |= 4;
          else if (fy <= fx + DELTA) fl |= 8;
        }
        for (int o = 16; o; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
        if (lane == 0) {
          if (fl & 1) vcl[x] = 1;
          else if ((fl & (2 | 8)) || !(fl & 4)) bad = 1;
          else vcl[x] = 2;
        }
      }
    }
    bad = __syncthreads_or(bad);
    if (!bad) {
      // grounding of the B vertices along exactly equal neighbours
      const int ncand = s_ncand;
      for (int round = 0; round <= ncand; round++) {
        int ch = 0, left = 0;
        for (int i = wid; i < ncand; i += nw) {
          const int x = cand[i];
          if (vcl[x] != 2) continue;
          const double fx = fval[x];
          const int a = astart0[x], dg = adeg0[x];
          int g1 = 0;
          for (int j = lane; j < dg && !g1; j += 32) {
            const int y = bitmap_rank(gbm0, c.W, gcol0[a + j]);
            if (y >= 0 && fval[y] == fx && vcl[y] == 1) g1 = 1;
          }
          g1 = __any_sync(0xffffffffu, g1);
          if (g1) { if (lane == 0) vcl[x] = 1; ch = 1; } else left = 1;
        }
        const int anych = __syncthreads_or(ch);
        const int anyleft = __syncthreads_or(left);
        if (!anyleft) break;
        if (!anych) { bad = 1; break; }
      }
    }
    bad = __syncthreads_or(bad);
    if (!bad) {
      const unsigned long long km = s_fmax;
      for (int x = tid; x < n; x += nt) if (f64_to_ordered(fval[x]) == km) atomicMin(&s_arg, x);
      __syncthreads();
      if (tid == 0) {
        const int lmin = min(lu, lv), lmax = s_arg;
        const int64_t po = c.poff(t);
        c.tnb[t] = -1;  // kernel 3v: nothing to sweep
        c.tminv[t] = lmin; c.tmaxv[t] = lmax;
        c.pkind[po] = TLC_K_ESS;               // [min_value, max_value]   accelerated_PD.py:110
        c.pbv[po] = lmin; c.pdv[po] = lmax;
        c.pbirth[po] = fval[lmin]; c.pdeath[po] = fval[lmax];
        c.tnp[t] = 1;
        c.tnneg[t] = 0; c.tnpos[t] = 0;
      }
      return;
    }
    __syncthreads();
  }

  // ---- 1. vertex order ----
  // 32-bit pass: stable LSD radix sort on the order-preserving image of the value rounded DOWN to float (monotone,
  // so only vertices sharing a float can be out of place); then every run of equal floats is put into exact
  // (float64 value, id) order by the thread that finds its start -- runs are exact ties (already in id order by
  // stability) or a handful of near-equal values.  Half the radix passes of a 64-bit sort.
  // the sort's ping-pong buffers (keys + payload, 16 B per vertex) live in shared memory when the launch made room for the
  // sub-range's largest vicinity (sort_cap >= n): every radix pass then scatters through shared memory instead of the arena
  uint32_t* sort_base = reinterpret_cast<uint32_t*>(dyn + smem_ints) + (bm_in_smem ? 2 * c.W : 0);
  const bool sort_smem = sort_cap >= n && sort_cap > 0;
  uint32_t* k0 = sort_smem ? sort_base : reinterpret_cast<uint32_t*>(c.v64a + vo);
  uint32_t* k1 = sort_smem ? sort_base + sort_cap : reinterpret_cast<uint32_t*>(c.v64b + vo);
  uint32_t* p0 = sort_smem ? sort_base + 2 * sort_cap : reinterpret_cast<uint32_t*>(c.vs0 + vo);
  uint32_t* p1 = sort_smem ? sort_base + 3 * sort_cap : reinterpret_cast<uint32_t*>(c.vs1 + vo);
  // normalised filtrations live in [0, 1]: a 24-bit fixed-point image floor(f * 2^24) is monotone too and saves a radix
  // pass; collisions (values closer than 6e-8) are equal-key runs, fixed below exactly like equal floats
  // (the strided passes of this kernel stage four independent loads per thread before they use them: a vicinity has ~10
  //  vertices per thread and every pass would otherwise pay one memory latency per vertex)
  int out_of_unit = 0;
  for (int x0 = tid; x0 < n; x0 += 4 * nt) {
    double f4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) f4[u] = x0 + u * nt < n ? fval[x0 + u * nt] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++) out_of_unit |= !(f4[u] >= 0.0 && f4[u] <= 1.0);
  }
  const bool unit = __syncthreads_or(out_of_unit) == 0;
  for (int x0 = tid; x0 < n; x0 += 4 * nt) {
    double f4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) f4[u] = x0 + u * nt < n ? fval[x0 + u * nt] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int x = x0 + u * nt;
      if (x >= n) continue;
      if (unit) {
        const double q = floor(f4[u] * 16777216.0);
        k0[x] = q >= 16777215.0 ? 16777215u : (uint32_t)q;
      } else {
        const uint32_t b = (uint32_t)__float_as_int(__double2float_rd(f4[u]));
        k0[x] = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
      }
      p0[x] = x;
    }
  }
  __syncthreads();
  const int r = block_radix_sort<uint32_t>(k0, p0, k1, p1, n, unit ? 24 : 32, sh);
  const uint32_t* kf = r ? k1 : k0;
  uint32_t* ps = r ? p1 : p0;
  int32_t* need = c.vs2 + vo;  // need[s] = 1: the run of equal floats starting at s holds distinct float64 values
  for (int i = tid; i < n; i += nt) need[i] = 0;
  __syncthreads();
  for (int i = tid + 1; i < n; i += nt) {
    if (kf[i] != kf[i - 1] || f64_to_ordered(fval[ps[i]]) == f64_to_ordered(fval[ps[i - 1]])) continue;
    int s0 = i - 1;
    while (s0 > 0 && kf[s0 - 1] == kf[i]) s0--;
    need[s0] = 1;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    if (!need[i]) continue;  // exact ties stay as the stable sort left them: ascending id
    int j = i + 1;
    while (j < n && kf[j] == kf[i]) j++;
    for (int a = i + 1; a < j; a++) {  // insertion sort by (value, id)
      const uint32_t xa = ps[a];
      const unsigned long long ka = f64_to_ordered(fval[xa]);
      int b = a - 1;
      while (b >= i) {
        const uint32_t xb = ps[b];
        const unsigned long long kb = f64_to_ordered(fval[xb]);
        if (kb < ka || (kb == ka && xb < xa)) break;
        ps[b + 1] = xb;
        b--;
      }
      ps[b + 1] = xa;
    }
  }
  __syncthreads();
  // exact keys in sorted order (block cuts, distinct-value flags)
  unsigned long long* ks = (r ? c.v64a : c.v64b) + vo;  // the key buffer not holding kf
  for (int i0 = tid; i0 < n; i0 += 4 * nt) {
    uint32_t x4[4];
    double f4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) x4[u] = i0 + u * nt < n ? ps[i0 + u * nt] : 0u;
#pragma unroll
    for (int u = 0; u < 4; u++) f4[u] = i0 + u * nt < n ? fval[x4[u]] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++) if (i0 + u * nt < n) ks[i0 + u * nt] = f64_to_ordered(f4[u]);
  }
  __syncthreads();
  int32_t* flag = c.vs2 + vo;
  const double fmin = fval[ps[0]];
  const double pmin = __dmul_rn(__dadd_rn(fmin, 1.0), 1e-6);
  for (int i0 = tid; i0 < n; i0 += 4 * nt) {
    uint32_t x4[4];
    unsigned long long kp4[4], kx4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * nt;
      x4[u] = i < n ? ps[i] : 0u;
      kx4[u] = i < n ? ks[i] : 0ull;
      kp4[u] = (i < n && i > 0) ? ks[i - 1] : 0ull;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * nt;
      if (i >= n) continue;
      const int32_t x = (int32_t)x4[u];
      vord[i] = x;
      vrank[x] = i;
      int f = 1;
      if (i > 0) {
        // same block as the predecessor unless every key owned by it is strictly below every key owned here
        // (the exact values in sorted order were just written to ks[] as ordered bit patterns: decode instead of gathering)
        const double fp = ordered_to_f64(kp4[u]), fx = ordered_to_f64(kx4[u]);
        const double hi_prev = __dadd_rn(fp, __dmul_rn(__dadd_rn(fp, 1.0), 1e-6));
        const double lo_here = __dadd_rn(fx, pmin);
        f = hi_prev < lo_here ? 1 : 0;
      }
      flag[i] = f;
    }
  }
  __syncthreads();
  // ---- 2. blocks ----
  int32_t* own = reinterpret_cast<int32_t*>(r ? p0 : p1);  // the payload buffer not holding the result
  for (int i0 = tid; i0 < n; i0 += 4 * nt) {
    int f4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) f4[u] = i0 + u * nt < n ? flag[i0 + u * nt] : 0;
#pragma unroll
    for (int u = 0; u < 4; u++) if (i0 + u * nt < n) own[i0 + u * nt] = f4[u];
  }
  __syncthreads();
  const int nb = block_exclusive_scan(flag, n, sh.scan);  // flag[i] = #starts before i
  for (int i = tid; i < n; i += nt) if (own[i]) bfirst[flag[i]] = i | 0x40000000;  // bit 30 cleared below where it fails
  if (tid == 0) { bfirst[nb] = n; c.tnb[t] = nb; c.tminv[t] = (int32_t)ps[0]; }
  __syncthreads();
  // block index of every LOCAL id, in shared memory when it fits (16-bit when nb < 65536)
  const bool in_smem = n <= smem_ints;
  int32_t* sblk = in_smem ? dyn : c.vcls + vo;
  const unsigned long long klast = ks[n - 1];
  for (int i0 = tid; i0 < n; i0 += 4 * nt) {
    int fl4[4], ow4[4];
    uint32_t x4[4];
    unsigned long long kp4[4], kx4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * nt;
      const bool ok = i < n;
      fl4[u] = ok ? flag[i] : 0; ow4[u] = ok ? own[i] : 0; x4[u] = ok ? ps[i] : 0u;
      kx4[u] = ok ? ks[i] : 0ull; kp4[u] = (ok && i > 0) ? ks[i - 1] : 0ull;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * nt;
      if (i >= n) continue;
      const int bidx = fl4[u] - 1 + ow4[u];
      sblk[x4[u]] = bidx;
      if (i > 0 && !ow4[u] && kx4[u] != kp4[u]) atomicOr(&bfirst[bidx], (int32_t)0x80000000);
      // essential maximum: first rank attaining the largest value
      if (kx4[u] == klast && (i == 0 || kp4[u] != klast)) c.tmaxv[t] = (int32_t)x4[u];
    }
  }
  __syncthreads();
  const int32_t* __restrict__ astart = c.astart + vo;
  const int32_t* __restrict__ adeg = c.adeg + vo;
  const int32_t* __restrict__ hint = c.neg + vo;  // kernel 1b's tree parents (-1: none): towards the last root ...
  const int32_t* __restrict__ hint0 = in_smem ? c.vcls + vo : nullptr;  // ... and the first (c.vcls doubles as the block table when it does not fit shared memory)
  if (c.dbm) {
    // graph-row route: the row is the graph's CSR row, entries outside the vicinity bitmap are skipped
    // the bitmap and its word-prefix ranks staged behind the block table (shared memory) when the launch made room
    const uint32_t* gbm = c.dbm + (size_t)t * 2 * c.W;
    uint32_t* sbm = reinterpret_cast<uint32_t*>(dyn + smem_ints);
    if (bm_in_smem) {
      for (int w = tid; w < 2 * c.W; w += nt) sbm[w] = gbm[w];
      __syncthreads();
    }
    const uint32_t* __restrict__ bm = bm_in_smem ? sbm : gbm;
    const int32_t* __restrict__ gcol = c.gcol;
    for (int x0 = tid; x0 < n; x0 += 4 * nt) {
      int a4[4], dg4[4], h4[4], h04[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {  // (the four vertices' row bounds and tree-parent hints in flight together)
        const int x = x0 + u * nt;
        const bool ok = x < n;
        a4[u] = ok ? astart[x] : 0; dg4[u] = ok ? adeg[x] : 0;
        h4[u] = ok ? hint[x] : -1; h04[u] = (ok && hint0) ? hint0[x] : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int x = x0 + u * nt;
        if (x >= n) continue;
        const int bx = sblk[x];
        const int a = a4[u], dg = dg4[u];
        const int h = h4[u], h0 = h04[u];  // kernel 1b / 1t's tree parents: usually already in an earlier block
        bool has = (h >= 0 && sblk[h] < bx) || (h0 >= 0 && sblk[h0] < bx);
        for (int j = 0; j < dg && !has && bx > 0; j++) {
          const int y = bitmap_rank(bm, c.W, gcol[a + j]);
          has = y >= 0 && sblk[y] < bx;
        }
        if (!has) atomicAnd(&bfirst[bx], (int32_t)~0x40000000);
      }
    }
    return;
  }
  const uint32_t* __restrict__ anb = c.anb + ao;
  for (int x = tid; x < n; x += nt) {  // x: local id
    const int bx = sblk[x];
    const int a = astart[x], dg = adeg[x];
    const int h = hint[x], h0 = hint0 ? hint0[x] : -1;
    bool has = (h >= 0 && sblk[h] < bx) || (h0 >= 0 && sblk[h0] < bx);
    for (int j = 0; j < dg && !has && bx > 0; j++) has = sblk[anb[a + j]] < bx;  // (block 0 has no earlier block)
    if (!has) atomicAnd(&bfirst[bx], (int32_t)~0x40000000);
  }
}

}  // namespace

void launch_vorder(const Params& p, const ChunkView& c, int t0, int cnt, int block, int64_t n_max, cudaStream_t st) {
  const int smem_ints = n_max * 4 <= 160 * 1024 ? (int)n_max : 0;
  const int bm_in_smem = c.dbm != nullptr && (size_t)c.W * 8 <= 48 * 1024;
  size_t bytes = (size_t)smem_ints * 4 + (bm_in_smem ? (size_t)c.W * 8 : 0);
  int sort_cap = (int)((n_max + 3) / 4 * 4);
  // (measured on B200, Computers-shaped 2-hop, 4096 targets per step: 3.19 ms with the sort buffers in shared memory vs 2.41 ms
  //  through the arena -- the 16 B per vertex cost a resident CTA per SM, and the sort is latency-bound either way: off by default)
  if (smem_ints == 0 || bytes + (size_t)sort_cap * 16 > 200 * 1024 || !getenv("TLC_VORDER_SMEM_SORT")) sort_cap = 0;
  bytes += (size_t)sort_cap * 16;
  cudaFuncSetAttribute((const void*)vorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  cudaFuncSetAttribute((const void*)vorder_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (const char* env = getenv("TLC_VORDER_BLOCK")) block = atoi(env);  // (tuning experiments)
  ChunkView c2 = c;
  c2.no_fast = getenv("TLC_NO_FAST_DIAGRAM") ? 1 : 0;  // (parity experiments: every target through the sort and the sweep)
  vorder_kernel<<<cnt, block, bytes, st>>>(p, c2, t0, smem_ints, bm_in_smem, sort_cap);
  count_launch();
}

}  // namespace tlc
