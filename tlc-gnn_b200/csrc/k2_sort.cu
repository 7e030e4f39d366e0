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
#include "tlc_sort.cuh"

namespace tlc {
namespace {

__global__ void sort_kernel(Params p, ChunkView c, int sweep_mask, int fb_only) {
  __shared__ SortShared sh;
  __shared__ int sh_bad;
  const int t = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = c.tn[t], m = c.tm[t];
  if (n == 0 || c.tstatus[t] > TLC_ST_TRIVIAL) return;
  if (fb_only && !c.tfb[t]) sweep_mask &= ~1;  // kernel 3v finished this target's ascending sweep
  if (!sweep_mask) return;
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
    if (!((sweep_mask >> sweep) & 1)) continue;
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

void launch_sort(const Params& p, const ChunkView& c, int block, int sweep_mask, int fb_only, cudaStream_t st) {
  sort_kernel<<<c.T, block, 0, st>>>(p, c, sweep_mask, fb_only);
  count_launch();
}

}  // namespace tlc
