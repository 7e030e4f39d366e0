// k2v_vorder.cu -- kernel 2v: the vertex order of every vicinity and its rank-space lower adjacency.
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
//   1. stable LSD radix sort of the n vertices on the ordered image of their float64 value
//      (ties keep ascending local id)                    -> vord[rank] = local id, vrank[local id] = rank
//   2. block starts (bit 31 of bfirst[b]: the block holds distinct values, i.e. a near-tie block)
//   3. counting sort of the m edges by owner rank        -> loff[n+1], ladj[m] = rank of the other endpoint
//      (bit 31: it lies in an earlier block), bend[b] = end of block b's edges | block flags
//   4. local ids of the essential pair [min, max]        (accelerated_PD.py:35-38,110: first vertex in
//      ascending id attaining the extreme value)
#include "tlc_common.cuh"
#include "tlc_sort.cuh"

namespace tlc {
namespace {

__global__ void vorder_kernel(Params p, ChunkView c, int smem_ints) {
  extern __shared__ int32_t dyn[];
  __shared__ SortShared sh;
  const int t = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = c.tn[t], m = c.tm[t];
  if (tid == 0) c.tfb[t] = 0;
  if (n == 0 || c.tstatus[t] > TLC_ST_TRIVIAL) return;
  const int64_t vo = c.voff[t], eo = c.eoff[t];
  const double* __restrict__ fval = c.fval + vo;
  int32_t* vord = c.vord + vo;
  int32_t* vrank = c.vrank + vo;
  int32_t* bfirst = c.bfirst + vo + t;
  int32_t* loff = c.loff + vo + t;
  int32_t* bend = c.bend + vo + t;

  // ---- 1. vertex order ----
  unsigned long long* k0 = c.v64a + vo;
  unsigned long long* k1 = c.v64b + vo;
  uint32_t* p0 = reinterpret_cast<uint32_t*>(c.vs0 + vo);
  uint32_t* p1 = reinterpret_cast<uint32_t*>(c.vs1 + vo);
  for (int x = tid; x < n; x += nt) { k0[x] = f64_to_ordered(fval[x]); p0[x] = x; }
  __syncthreads();
  const int r = block_radix_sort<unsigned long long>(k0, p0, k1, p1, n, 64, sh);
  const unsigned long long* ks = r ? k1 : k0;
  const uint32_t* ps = r ? p1 : p0;
  int32_t* flag = c.vs2 + vo;
  const double fmin = fval[ps[0]];
  const double pmin = __dmul_rn(__dadd_rn(fmin, 1.0), 1e-6);
  for (int i = tid; i < n; i += nt) {
    const int32_t x = (int32_t)ps[i];
    vord[i] = x;
    vrank[x] = i;
    int f = 1;
    if (i > 0) {
      // same block as the predecessor unless every key owned by it is strictly below every key owned here
      const double fp = fval[ps[i - 1]], fx = fval[x];
      const double hi_prev = __dadd_rn(fp, __dmul_rn(__dadd_rn(fp, 1.0), 1e-6));
      const double lo_here = __dadd_rn(fx, pmin);
      f = hi_prev < lo_here ? 1 : 0;
    }
    flag[i] = f;
  }
  __syncthreads();
  // ---- 2. blocks ----
  int32_t* own = reinterpret_cast<int32_t*>(r ? p0 : p1);  // the payload buffer not holding the result
  for (int i = tid; i < n; i += nt) own[i] = flag[i];
  __syncthreads();
  const int nb = block_exclusive_scan(flag, n, sh.scan);  // flag[i] = #starts before i
  for (int i = tid; i < n; i += nt) if (own[i]) bfirst[flag[i]] = i;
  if (tid == 0) { bfirst[nb] = n; c.tnb[t] = nb; c.tminv[t] = (int32_t)ps[0]; }
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    if (i > 0 && !own[i] && ks[i] != ks[i - 1]) atomicOr(&bfirst[flag[i] - 1 + own[i]], (int32_t)0x80000000);
    // essential maximum: first rank attaining the largest value
    if (ks[i] == ks[n - 1] && (i == 0 || ks[i - 1] != ks[n - 1])) c.tmaxv[t] = (int32_t)ps[i];
  }
  __syncthreads();
  // ---- 3. rank-space lower adjacency ----
  // srank[local id] = rank | block index << 16 (both < 65536 here)
  if (n >= 65536) {  // 16-bit packing does not hold: leave this target to the edge-sorted kernels 2 + 3
    if (tid == 0) c.tfb[t] = 1;
    return;
  }
  const bool in_smem = 2 * n <= smem_ints;
  int32_t* cnt = in_smem ? dyn : c.vs2 + vo;         // flag[] is dead from here on
  uint32_t* srank = in_smem ? reinterpret_cast<uint32_t*>(dyn + n) : reinterpret_cast<uint32_t*>(c.vcls + vo);
  for (int i = tid; i < n; i += nt) {
    cnt[i] = 0;
    srank[ps[i]] = (uint32_t)i | ((uint32_t)(flag[i] - 1 + own[i]) << 16);
  }
  __syncthreads();
  const int32_t* __restrict__ elo = c.elo + eo;
  const int32_t* __restrict__ ehi = c.ehi + eo;
  for (int e = tid; e < m; e += nt) {
    const int ra = (int)(srank[elo[e]] & 0xffffu), rb = (int)(srank[ehi[e]] & 0xffffu);
    atomicAdd(&cnt[max(ra, rb)], 1);
  }
  __syncthreads();
  block_exclusive_scan(cnt, n, sh.scan);
  for (int i = tid; i < n; i += nt) loff[i] = cnt[i];
  if (tid == 0) loff[n] = m;
  __syncthreads();
  // ladj entry = rank of the earlier endpoint | bit 31 when it lies in an EARLIER block than the owner
  uint32_t* ladj = c.ladj + eo;
  for (int e = tid; e < m; e += nt) {
    const uint32_t wa = srank[elo[e]], wb = srank[ehi[e]];
    const uint32_t wo = (wa & 0xffffu) > (wb & 0xffffu) ? wa : wb, wy = wo == wa ? wb : wa;
    const int pos = atomicAdd(&cnt[wo & 0xffffu], 1);
    ladj[pos] = (wy & 0xffffu) | ((wy >> 16) != (wo >> 16) ? 0x80000000u : 0u);
  }
  __syncthreads();  // ladj of this vicinity complete (global writes of the block visible to the block)
  // per block: end of its owned edges | bit 31 distinct values | bit 30 every vertex has an earlier-block neighbour
  for (int b = tid; b < nb; b += nt) {
    const int s0 = bfirst[b] & 0x7fffffff, s1 = bfirst[b + 1] & 0x7fffffff;
    bool all_out = true;
    for (int x = s0; x < s1 && all_out; x++) {
      bool has = false;
      for (int j = loff[x]; j < loff[x + 1] && !has; j++) has = (ladj[j] >> 31) != 0;
      all_out = has;
    }
    bend[b] = loff[s1] | (bfirst[b] & (int32_t)0x80000000) | (all_out ? 0x40000000 : 0);
  }
}

}  // namespace

void launch_vorder(const Params& p, const ChunkView& c, int block, int64_t n_max, cudaStream_t st) {
  const int64_t want = 2 * n_max;
  const int smem_ints = want * 4 <= 160 * 1024 ? (int)want : 0;
  const size_t bytes = (size_t)smem_ints * 4;
  if (bytes > 8 * 1024)
    cudaFuncSetAttribute((const void*)vorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  vorder_kernel<<<c.T, block, bytes, st>>>(p, c, smem_ints);
  count_launch();
}

}  // namespace tlc
