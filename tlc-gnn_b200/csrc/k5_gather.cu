// k5_gather.cu -- kernel 5: hand-off of the cached persistence-image table to the link-prediction decoder.
//
// The reference keeps pi_sg as a host float64[E, res^2] array and, on EVERY decode call, fancy-indexes the
// step's rows on the host, converts them to float32 and uploads them (baselines/TLCGNN.py:35-53:
// `PI = np.concatenate((self.PI[:train_pos], self.PI[train_pos:train_pos+train_neg][index]))`,
// `new_x = torch.Tensor(PI.reshape(len(total_edges), -1)).cuda()`).  Here the table stays resident in HBM
// (float64, the .npy cache layout of loaddatas.py:62-64,102) and the step's rows are gathered and rounded to
// float32 on the device: out[i, :] = (float) table[row(i), :], row(i) = index ? index[i] : start + i.
//
// HBM-bound: 8 B read + 4 B written per element, 8 B per index.  One thread per output element, consecutive
// threads walk consecutive elements of a row (coalesced reads within a row, fully coalesced writes).
#include <algorithm>

#include "tlc_common.cuh"

namespace tlc {
namespace {

__global__ void __launch_bounds__(256) gather_rows_kernel(const double* __restrict__ table, int64_t rows, int r2,
                                                          const int64_t* __restrict__ index, int64_t start, int64_t n,
                                                          float* __restrict__ out, int* __restrict__ bad) {
  const int64_t total = n * r2;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / r2;
    const int j = (int)(e - i * r2);
    const int64_t r = index ? index[i] : start + i;
    if (r < 0 || r >= rows) {  // numpy raises IndexError: reported to the host, the row is zero-filled
      if (j == 0) atomicExch(bad, 1);
      out[e] = 0.f;
      continue;
    }
    out[e] = (float)table[r * r2 + j];  // torch.Tensor(float64 ndarray): round to nearest float32
  }
}

// rows of a sub-call back to their place in the caller's tables: dst[idx[i], :] = src[i, :]
__global__ void __launch_bounds__(256) scatter_rows_kernel(const double* __restrict__ src_pi, const float* __restrict__ src_pi32,
                                                           const uint8_t* __restrict__ src_st, const int64_t* __restrict__ idx,
                                                           int64_t k, int r2, double* dst_pi, float* dst_pi32, uint8_t* dst_st) {
  const int64_t total = k * r2;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / r2;
    const int j = (int)(e - i * r2);
    const int64_t r = idx[i];
    dst_pi[r * r2 + j] = src_pi[e];
    if (dst_pi32) dst_pi32[r * r2 + j] = src_pi32[e];
    if (j == 0 && dst_st) dst_st[r] = src_st[i];
  }
}

// targets of a sub-list, and the rows they came from (kernel S hands rows it cannot take to the staged pipeline)
__global__ void gather_targets_kernel(const int32_t* __restrict__ targets, const int32_t* __restrict__ list, int64_t k,
                                      int32_t* __restrict__ sub, int64_t* __restrict__ idx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = list[i];
    sub[2 * i] = targets[2 * r];
    sub[2 * i + 1] = targets[2 * r + 1];
    idx[i] = r;
  }
}

// ---- multi-GPU exchange without a library collective (fused variant) ----
// Every rank owns a table [header | float32 rows[rows][r2 + 1]] in its HBM (25 image floats + the status as the last
// float); the tables of the peers are mapped through CUDA IPC.  This kernel stores the rows a rank has just computed AT
// THEIR FINAL INDEX into the table of EVERY rank -- peer stores over NVLink / NVSwitch for the remote ones -- so there
// is no padding copy, no all-gather and no un-permute pass.  The last block then publishes completion: a system-scope
// fence followed by one atomic increment of the arrival counter in every table's header.
__global__ void __launch_bounds__(256) peer_scatter_kernel(const float* __restrict__ src_pi32, const uint8_t* __restrict__ src_st,
                                                           const int64_t* __restrict__ row_index, int64_t k, int r2,
                                                           PeerTables pt, unsigned int* ticket) {
  const int w = r2 + 1;
  const int64_t total = k * w;
  for (int t = 0; t < pt.n; t++) {
    float* dst = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(pt.table[t]) + PEER_HEADER_BYTES);
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i = e / w;
      const int j = (int)(e - i * w);
      const int64_t r = row_index[i];
      dst[r * w + j] = j < r2 ? src_pi32[i * r2 + j] : (float)src_st[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(ticket, 1u);
    if (done == gridDim.x - 1) {  // every block's stores are fenced: signal all tables (this rank's own included)
      *ticket = 0;
      __threadfence_system();
      for (int t = 0; t < pt.n; t++) atomicAdd_system(reinterpret_cast<unsigned int*>(pt.table[t]), 1u);
    }
  }
}

// wait until `target` arrivals were counted in this rank's own table header (one per rank and exchange step)
__global__ void peer_wait_kernel(const unsigned int* flag, unsigned int target) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    while (*reinterpret_cast<const volatile unsigned int*>(flag) < target) __nanosleep(200);
    __threadfence_system();
  }
}

}  // namespace

void launch_peer_scatter(const float* src_pi32, const uint8_t* src_st, const int64_t* row_index, int64_t k, int r2,
                         const PeerTables& pt, unsigned int* ticket, int sm_count, cudaStream_t st) {
  const int64_t want = std::max<int64_t>((k * (r2 + 1) + 255) / 256, 1);
  const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count * 4);
  peer_scatter_kernel<<<grid, 256, 0, st>>>(src_pi32, src_st, row_index, k, r2, pt, ticket);
  count_launch();
}

void launch_peer_wait(const unsigned int* flag, unsigned int target, cudaStream_t st) {
  peer_wait_kernel<<<1, 32, 0, st>>>(flag, target);
  count_launch();
}

void launch_gather_targets(const int32_t* targets, const int32_t* list, int64_t k, int32_t* sub, int64_t* idx, cudaStream_t st) {
  if (k <= 0) return;
  const int grid = (int)std::min<int64_t>((k + 255) / 256, 2048);
  gather_targets_kernel<<<grid, 256, 0, st>>>(targets, list, k, sub, idx);
  count_launch();
}

void launch_scatter_rows(const double* src_pi, const float* src_pi32, const uint8_t* src_st, const int64_t* idx, int64_t k,
                         int r2, double* dst_pi, float* dst_pi32, uint8_t* dst_st, int sm_count, cudaStream_t st) {
  if (k <= 0) return;
  const int64_t want = (k * r2 + 255) / 256;
  const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count * 16);
  scatter_rows_kernel<<<grid, 256, 0, st>>>(src_pi, src_pi32, src_st, idx, k, r2, dst_pi, dst_pi32, dst_st);
  count_launch();
}

void launch_gather_rows(const double* table, int64_t rows, int r2, const int64_t* index, int64_t start, int64_t n,
                        float* out, int* bad, int sm_count, cudaStream_t st) {
  const int64_t total = n * r2;
  if (total <= 0) return;
  const int64_t want = (total + 255) / 256;
  const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count * 16);  // a multiple of the SM count, grid-stride
  gather_rows_kernel<<<grid, 256, 0, st>>>(table, rows, r2, index, start, n, out, bad);
  count_launch();
}

}  // namespace tlc
