// k1c_edgelist.cu -- kernel 1c: the canonical edge list of a vicinity, derived from its adjacency.
//
// The edge-sorted kernels 2 / 3 / 3b (descending sweep, Pos/Neg lists, loops, and the ascending sweep of
// targets kernel 3v hands back) and the diagram outputs work on the induced edges in the canonical order
// of the oracle: lexicographic (lo, hi) in local ids, oriented (lo, hi) -- the order the reference sees
// when its sub-graph is built with ascending node and edge insertion.  Kernel 1 writes the
// adjacency (rows ascending); this kernel expands its upper half: count per row, scan, fill.
// One CTA per vicinity; skipped for targets that never reach an edge-sorted kernel.
#include "tlc_common.cuh"

namespace tlc {
namespace {

__global__ void edgelist_kernel(Params p, ChunkView c, int fb_only) {
  __shared__ int32_t scan[1025];
  const int t = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = c.tn[t];
  if (n == 0) return;
  if (fb_only && (!c.tfb[t] || c.tstatus[t] > TLC_ST_TRIVIAL)) return;  // (all live targets when diagrams are wanted)
  const int64_t vo = c.voff[t], eo = c.eoff[t], ao = c.aoff[t];
  const int32_t* __restrict__ astart = c.astart + vo;
  const int32_t* __restrict__ adeg = c.adeg + vo;
  const uint32_t* __restrict__ anb = c.anb + ao;
  const double* __restrict__ aw = c.aw + ao;
  int32_t* cnt = c.vs0 + vo;
  // rows are ascending: the upper neighbours (y > x) are a suffix; find its start by bisection
  for (int x = tid; x < n; x += nt) {
    const int a = astart[x], dg = adeg[x];
    int lo = 0, hi = dg;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int)anb[a + mid] > x) hi = mid; else lo = mid + 1;
    }
    cnt[x] = dg - lo;
  }
  __syncthreads();
  block_exclusive_scan(cnt, n, scan);
  for (int x = tid; x < n; x += nt) {
    const int a = astart[x], dg = adeg[x];
    const int64_t o0 = eo + cnt[x];
    const int k = x + 1 < n ? cnt[x + 1] - cnt[x] : c.tm[t] - cnt[x];
    for (int i = 0; i < k; i++) {
      c.elo[o0 + i] = x;
      c.ehi[o0 + i] = (int32_t)anb[a + dg - k + i];
      c.ew[o0 + i] = aw[a + dg - k + i];
    }
  }
}

}  // namespace

void launch_edgelist(const Params& p, const ChunkView& c, int block, int fb_only, cudaStream_t st) {
  edgelist_kernel<<<c.T, block, 0, st>>>(p, c, fb_only);
  count_launch();
}

}  // namespace tlc
