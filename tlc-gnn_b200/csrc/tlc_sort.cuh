// tlc_sort.cuh -- block-wide stable LSD radix sort and the reference's perturbed float64 edge keys,
// shared by kernel 2 (edge orders) and kernel 2v (vertex order).
#pragma once
#include "tlc_common.cuh"

namespace tlc {

constexpr int SORT_MAX_WARPS = 16;  // blocks of at most 512 threads
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
    for (int i0 = tid; i0 < n; i0 += 4 * nt) {  // (four loads in flight per thread)
      K k4[4];
#pragma unroll
      for (int u = 0; u < 4; u++) k4[u] = i0 + u * nt < n ? kin[i0 + u * nt] : (K)0;
#pragma unroll
      for (int u = 0; u < 4; u++) if (i0 + u * nt < n) atomicAdd(&sh.hist[digit_of(k4[u], shift)], 1);
    }
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
      uint32_t pay[SORT_ITEMS];
      int rnk[SORT_ITEMS];
      const int wbase = t0 + wid * 32 * SORT_ITEMS;
#pragma unroll
      for (int it = 0; it < SORT_ITEMS; it++) {  // (keys and payloads of the tile in flight together)
        const int i = wbase + it * 32 + lane;
        key[it] = i < n ? kin[i] : (K)0;
        pay[it] = i < n ? pin[i] : 0u;
      }
#pragma unroll
      for (int it = 0; it < SORT_ITEMS; it++) {
        const int i = wbase + it * 32 + lane;
        const bool ok = i < n;
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
          pout[pos] = pay[it];
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


}  // namespace tlc
