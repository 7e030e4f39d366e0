// k2_sort.cu -- kernel 2: the two simplex orders of every vicinity (segmented stable sort).
//
// Replaces perturb_filter_function (accelerated_PD.py:6-23) and the two `simplices.sort(...)` calls of
// Union_find (accelerated_PD.py:40-41 ascending, :76-77 descending).  Vertices always precede their
// incident edges in either sweep (asc > max endpoint, desc < min endpoint), so only the EDGE order
// matters; it is returned as two permutations of the canonical (lexicographic) edge list:
//   ord_asc : ascending  (asc key,  canonical index)      == python's stable sort by (value, len)
//   ord_desc: descending desc key, ties in canonical index == stable sort(reverse=True)
// with the reference's float64 keys  asc = max + (min + 1)*1e-6,  desc = min - (101 - max)*1e-6
// evaluated in exactly that parenthesisation without FMA contraction.
//
// Fast path: vertices are ranked by filtration value into dense classes (equal value <=> equal
// class); the composite integer key (class(max) , class(min)) has at most 2*ceil(log2 C) bits and is
// sorted with a stable LSD radix sort (8-bit digits, digits that are constant over the segment are
// skipped) -- 1 pass for hop-distance filtrations, <= 4 for n <= 65536 distinct values.  The result is
// then VERIFIED against the float64 keys (adjacent pairs, parallel); only when the perturbed keys
// cross the lexicographic order (vertex values closer than ~1e-6, F4) or the float keys tie where
// the classes do not, the segment is re-sorted on the 64-bit ordered image of the float keys.
#include "tlc_common.cuh"

namespace tlc {
namespace {

constexpr int SORT_MAX_WARPS = 32;
constexpr int SORT_ITEMS = 4;

struct SortShared {
  int32_t hist[256];
  int32_t base[256];
  int32_t whist[SORT_MAX_WARPS][256];
  int32_t scan[1025];
  int32_t flag;
};

template <typename K>
__device__ __forceinline__ int digit_of(K k, int shift) { return (int)((k >> shift) & (K)255); }

// stable LSD radix sort of (key, payload) pairs living in global memory; ping-pong between (k0,p0) and
// (k1,p1).  Returns 0 if the result is in (k0,p0), 1 if in (k1,p1).  All threads of the block call.
template <typename K>
__device__ int block_radix_sort(K* k0, uint32_t* p0, K* k1, uint32_t* p1, int n, int key_bits, SortShared& sh) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  int cur = 0;
  for (int shift = 0; shift < key_bits; shift += 8) {
    K* kin = cur ? k1 : k0;
    uint32_t* pin = cur ? p1 : p0;
    K* kout = cur ? k0 : k1;
    uint32_t* pout = cur ? p0 : p1;
    for (int d = tid; d < 256; d += nt) sh.hist[d] = 0;
    if (tid == 0) sh.flag = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) atomicAdd(&sh.hist[digit_of(kin[i], shift)], 1);
    __syncthreads();
    for (int d = tid; d < 256; d += nt) if (sh.hist[d] == n) sh.flag = 1;  // digit constant: skip the pass
    __syncthreads();
    if (sh.flag) { __syncthreads(); continue; }
    if (wid == 0) {  // exclusive scan of the 256 counts by one warp (8 per lane)
      int loc[8], s = 0;
      for (int j = 0; j < 8; j++) { loc[j] = sh.hist[lane * 8 + j]; s += loc[j]; }
      int inc = s;
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
      int run = inc - s;
      for (int j = 0; j < 8; j++) { sh.base[lane * 8 + j] = run; run += loc[j]; }
    }
    __syncthreads();
    const int tile = nt * SORT_ITEMS;
    for (int t0 = 0; t0 < n; t0 += tile) {
      for (int i = tid; i < nw * 256; i += nt) (&sh.whist[0][0])[i] = 0;
      __syncthreads();
      K key[SORT_ITEMS];
      int rnk[SORT_ITEMS];
      const int wbase = t0 + wid * 32 * SORT_ITEMS;
#pragma unroll
      for (int it = 0; it < SORT_ITEMS; it++) {
        const int i = wbase + it * 32 + lane;
        const bool ok = i < n;
        key[it] = ok ? kin[i] : (K)0;
        const int d = ok ? digit_of(key[it], shift) : (256 + lane);  // inactive lanes never match
        const unsigned mask = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(mask) - 1;
        int old = 0;
        if (ok && lane == leader) { old = sh.whist[wid][d]; sh.whist[wid][d] = old + __popc(mask); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rnk[it] = old + __popc(mask & lanemask_lt());
        __syncwarp();
      }
      __syncthreads();
      for (int d = tid; d < 256; d += nt) {  // per digit: exclusive scan over the warps, on top of base
        int run = sh.base[d];
        for (int w = 0; w < nw; w++) { const int c = sh.whist[w][d]; sh.whist[w][d] = run; run += c; }
        sh.base[d] = run;
      }
      __syncthreads();
#pragma unroll
      for (int it = 0; it < SORT_ITEMS; it++) {
        const int i = wbase + it * 32 + lane;
        if (i < n) {
          const int pos = sh.whist[wid][digit_of(key[it], shift)] + rnk[it];
          kout[pos] = key[it];
          pout[pos] = pin[i];
        }
      }
      __syncthreads();
    }
    cur ^= 1;
  }
  return cur;
}

// the reference's perturbed keys -- accelerated_PD.py:18-21 -- IEEE double, this operation order
__device__ __forceinline__ double key_asc(double fa, double fb) {
  const double mx = fmax(fa, fb), mn = fmin(fa, fb);
  return __dadd_rn(mx, __dmul_rn(__dadd_rn(mn, 1.0), 1e-6));
}
__device__ __forceinline__ double key_desc(double fa, double fb) {
  const double mx = fmax(fa, fb), mn = fmin(fa, fb);
  return __dadd_rn(mn, -__dmul_rn(__dadd_rn(101.0, -mx), 1e-6));
}

__device__ __forceinline__ int bits_for(int c) {  // bits to represent 0..c-1, at least 1
  int b = 1;
  while ((1 << b) < c) b++;
  return b;
}

__global__ void sort_kernel(Params p, ChunkView c) {
  __shared__ SortShared sh;
  __shared__ int sh_bad;
  const int t = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = c.tn[t], m = c.tm[t];
  if (n == 0 || c.tstatus[t] > TLC_ST_TRIVIAL) return;
  const int64_t vo = c.voff[t], eo = c.eoff[t];
  const double* __restrict__ fval = c.fval + vo;
  const int32_t* __restrict__ elo = c.elo + eo;
  const int32_t* __restrict__ ehi = c.ehi + eo;
  int32_t* vcls = c.vcls + vo;

  // ---- vertex classes: sort vertices by value, dense-rank distinct values ----
  {
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
    for (int i = tid; i < n; i += nt) flag[i] = (i > 0 && ks[i] != ks[i - 1]) ? 1 : 0;
    __syncthreads();
    // inclusive rank = exclusive scan + own flag
    int32_t* own = reinterpret_cast<int32_t*>(r ? p0 : p1);  // the other payload buffer is free now
    for (int i = tid; i < n; i += nt) own[i] = flag[i];
    __syncthreads();
    const int total = block_exclusive_scan(flag, n, sh.scan);
    for (int i = tid; i < n; i += nt) vcls[ps[i]] = flag[i] + own[i];
    if (tid == 0) c.tncls[t] = total + 1;
    __syncthreads();
  }
  if (m == 0) return;
  const int ncls = c.tncls[t];
  const int cb = bits_for(ncls);

  for (int sweep = 0; sweep < 2; sweep++) {
    uint32_t* ord = (sweep == 0 ? c.ord_asc : c.ord_desc) + eo;
    bool need_full = cb > 16;
    if (!need_full) {
      uint32_t* k0 = reinterpret_cast<uint32_t*>(c.sk0 + eo);
      uint32_t* k1 = reinterpret_cast<uint32_t*>(c.sk1 + eo);
      uint32_t* p0 = c.sp0 + eo;
      uint32_t* p1 = c.sp1 + eo;
      for (int e = tid; e < m; e += nt) {
        const int ca = vcls[elo[e]], cbb = vcls[ehi[e]];
        const int cmx = max(ca, cbb), cmn = min(ca, cbb);
        // ascending: (class(max), class(min)); descending: min descending then max descending
        k0[e] = sweep == 0 ? ((uint32_t)cmx << cb) | (uint32_t)cmn
                           : ((uint32_t)(ncls - 1 - cmn) << cb) | (uint32_t)(ncls - 1 - cmx);
        p0[e] = e;
      }
      __syncthreads();
      const int r = block_radix_sort<uint32_t>(k0, p0, k1, p1, m, 2 * cb, sh);
      const uint32_t* ps = r ? p1 : p0;
      // verify against the float64 keys (and copy out)
      if (tid == 0) sh_bad = 0;
      __syncthreads();
      int bad = 0;
      for (int i = tid; i < m; i += nt) {
        const uint32_t e1 = ps[i];
        ord[i] = e1;
        if (i > 0) {
          const uint32_t e0 = ps[i - 1];
          const double f0a = fval[elo[e0]], f0b = fval[ehi[e0]], f1a = fval[elo[e1]], f1b = fval[ehi[e1]];
          if (sweep == 0) {
            const double a0 = key_asc(f0a, f0b), a1 = key_asc(f1a, f1b);
            bad |= !(a0 < a1 || (a0 == a1 && e0 < e1));
          } else {
            const double a0 = key_desc(f0a, f0b), a1 = key_desc(f1a, f1b);
            bad |= !(a0 > a1 || (a0 == a1 && e0 < e1));
          }
        }
      }
      if (bad) sh_bad = 1;
      __syncthreads();
      need_full = sh_bad != 0;
      __syncthreads();
    }
    if (need_full) {
      unsigned long long* k0 = c.sk0 + eo;
      unsigned long long* k1 = c.sk1 + eo;
      uint32_t* p0 = c.sp0 + eo;
      uint32_t* p1 = c.sp1 + eo;
      for (int e = tid; e < m; e += nt) {
        const double fa = fval[elo[e]], fb = fval[ehi[e]];
        k0[e] = sweep == 0 ? f64_to_ordered(key_asc(fa, fb)) : ~f64_to_ordered(key_desc(fa, fb));
        p0[e] = e;
      }
      __syncthreads();
      const int r = block_radix_sort<unsigned long long>(k0, p0, k1, p1, m, 64, sh);
      const uint32_t* ps = r ? p1 : p0;
      for (int i = tid; i < m; i += nt) ord[i] = ps[i];
      __syncthreads();
    }
  }
}

}  // namespace

void launch_sort(const Params& p, const ChunkView& c, int block, cudaStream_t st) {
  sort_kernel<<<c.T, block, 0, st>>>(p, c);
  count_launch();
}

}  // namespace tlc
