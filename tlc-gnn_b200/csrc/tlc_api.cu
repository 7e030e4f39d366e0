// tlc_api.cu -- host orchestration and the C-ABI of libtlc_b200.so (include/tlc_b200.h).
//
// A call processes its targets in CHUNKS.  The counting pass of kernel 1 gives every target's vicinity size; the host
// orders the live targets (route, shared-memory class, size: largest first), packs as many as fit the HBM arena, and
// runs the stage kernels over the chunk.  Two routes (chosen per call, same results bit for bit):
//   materialised : 1 fill -> 1b filtration -> 2v vertex order -> 3v sweep -> [1c edge list -> 2 sort -> 3 union-find
//                  (descending sweep, Pos/Neg lists, targets 3v handed back)] -> [3b loops] -> 4 image
//   graph-row    : (light counting pass) -> 1b filtration on the graph's own CSR rows -> 2v -> 3v -> 4 image
//                  -- ascending-sweep-only calls on vicinities that are dense in the graph; no adjacency in HBM
// Per-target segments of every array are addressed through exclusive offsets uploaded per chunk; the image kernel
// scatters rows to their final position, so results keep the caller's target order.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "tlc_common.cuh"

namespace tlc {
static std::atomic_llong g_launches{0};
int64_t launch_count() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1); }
}  // namespace tlc

using namespace tlc;

static thread_local std::string g_err;
static int fail(int rc, const std::string& msg) {
  g_err = msg;
  return rc;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(TLC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + \
                                  std::to_string(__LINE__));                                       \
  } while (0)

// device allocation freed on every exit path (the CK macro returns early)
struct DevBuf {
  void* p = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  template <typename T_> T_* as() const { return static_cast<T_*>(p); }
};

static constexpr size_t ALIGN = 256;
static size_t align_up(size_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }

// bytes of arena per vertex / edge / pair slot (see carve())
static constexpr size_t BV = 4 * 6 + 8 * 3 + 8 * 3 + 4 * 6;           // vert vcls vs0 vs1 vs2 neg | fval d1 d2 | v64a v64b v64c | vord vrank bfirst astart adeg aminw
static constexpr size_t BE = 4 * 4 + 8 + 4 * 4 + 8 * 2 + 1;           // elo ehi pos arank | ew | ord_asc ord_desc sp0 sp1 | sk0 sk1 | isneg
static constexpr size_t BA = 4 + 8;                                   // anb | aw, per adjacency slot
static constexpr size_t BP = 1 + 4 * 2 + 8 * 2;                       // pkind | pbv pdv | pbirth pdeath
static constexpr size_t BT = 8 * 4 + 4 * 11 + 2 + 4;                  // tidx voff eoff aoff | tn tm tlu tlv tnp tnpos tnneg tncls tnb tminv tmaxv | tstatus tfb | bfirst terminator

struct tlc_graph {
  int device = 0;
  GraphView gv{};
  cudaStream_t stream = nullptr, own_stream = nullptr;
  // the filtration launches of a chunk's sub-ranges are independent: they run on side streams, forked from / joined to
  // the call's stream with events, so that the tail wave of one sub-range is filled by CTAs of the next
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  int64_t last_live = 0, last_nv = 0, last_ne = 0, last_fb = 0, last_general = 0, last_rowcheck = 0, last_blocks = 0;
  char* arena = nullptr;
  size_t arena_bytes = 0, arena_req = 0;
  // vicinity scratch (depends on hop)
  uint32_t* bitmaps = nullptr;
  int32_t* queue = nullptr;
  // ball cache (valid for vic_hop)
  uint32_t* ball_cache = nullptr;
  unsigned long long* ball_acc = nullptr;
  int32_t *ball_state = nullptr, *ball_list = nullptr;
  int vic_grid = 0, vic_hop = -1;
  int* work_counter = nullptr;
  ChunkView detail_chunk{};  // the chunk of the last tlc_vicinity_detail call (still carved in the arena)
  float* gminw = nullptr;  // [N] smallest kappa + 1 of each node's row, rounded down (graph-row route's settling margin)
  int64_t last_direct = 0;  // targets of the last call that took the graph-row route
  // density probe of the graph-row route (per hop / mode): decision and calls since it was taken
  int probe_hop = -1, probe_mode = -1, probe_age = 0;
  bool probe_direct = false;
  // per-call device buffers
  int32_t *d_n = nullptr, *d_m = nullptr, *d_ds = nullptr;
  uint8_t* d_st = nullptr;
  double* d_bytes = nullptr;
  int64_t call_cap = 0;
  // device staging of the host-buffer entry point (tlc_vicinity_pi): targets in, image rows + status out
  int32_t* io_t = nullptr;
  double* io_pi = nullptr;
  uint8_t* io_st = nullptr;
  int64_t io_cap_t = 0, io_cap_pi = 0;
  // host pinned staging
  char* h_pin = nullptr;
  size_t h_pin_bytes = 0;
  // stats of the last call
  double stage_ms[10] = {0};
  int nchunks = 0;
  double alg_bytes = 0, alg_bytes_bfs = 0, alg_bytes_uf = 0;
  bool timing = false;
  // kernel S (fused small-vicinity path): deferral lists, counters + statistics, their pinned host mirror
  int32_t *sm_list_b = nullptr, *sm_list_c = nullptr, *sm_list_big = nullptr, *sm_list_b2 = nullptr;
  cudaStream_t small_stream2 = nullptr;  // ... class C beside class B
  cudaEvent_t ev_small_c = nullptr;
  cudaStream_t small_stream = nullptr;   // classes B / C of kernel S run here while the staged pipeline takes the big targets
  cudaEvent_t ev_small_a = nullptr, ev_small_bc = nullptr;
  int32_t* sm_sub = nullptr;       // [cap][2] targets handed on to the staged pipeline
  int64_t* sm_idx = nullptr;       // [cap] their rows
  double* sm_pi = nullptr;         // [cap][r2] staged results of those rows, scattered back
  float* sm_pi32 = nullptr;
  uint8_t* sm_st = nullptr;
  int64_t sm_cap = 0, sm_sub_cap = 0;
  char* sm_dev = nullptr;          // [int counters[2] | pad | SmallStats]
  char* sm_host = nullptr;         // pinned mirror
  int small_skip = 0;              // calls left for which the small path is not tried (it handled too few rows)
  int small_hop = -1, small_mode = -1;
  double small_ms[3] = {0, 0, 0};
  int64_t small_rows[4] = {0, 0, 0, 0};  // last call: rows finished by class A / B / C / handed to the staged pipeline
  double hks_time = 0.1;   // diffusion time of TLC_F_FILT_HKS (data_utils_NC.py: hks_time = 0.1)
  // kernel 1t: per-root shortest-path tables of the whole graph (valid for sssp_plain), scratch of the build kernel
  SsspTables sssp{};
  double* sssp_pw = nullptr;
  int sssp_plain = -1;
  bool sssp_tried = false, sssp_all = false;
  double sssp_build_ms = 0;   // device time of the (one-time) table build
  int64_t last_table = 0;  // targets of the last call whose filtration came from the tables
  // multi-GPU exchange by peer stores: this rank's table, the mapped tables of all ranks, exchange epoch
  void* peer_own = nullptr;
  int64_t peer_rows = 0;
  int peer_r2 = 0;
  PeerTables peers{};
  std::vector<void*> peer_opened;   // tables mapped with cudaIpcOpenMemHandle
  unsigned int peer_epoch = 0;
  unsigned int* peer_ticket = nullptr;
  float* px_pi32 = nullptr;         // [cap][r2] staging of the shard's rows
  double* px_pi = nullptr;
  uint8_t* px_st = nullptr;
  int64_t px_cap = 0;
  std::vector<cudaEvent_t> ev_pool;  // stage-timing events, reused from call to call
  size_t ev_base = 0;                // first pool slot a (nested) call may use
  int sm_count = 0;
};

static int ensure_pinned(tlc_graph* g, size_t bytes) {
  if (bytes <= g->h_pin_bytes) return TLC_OK;
  if (g->h_pin) cudaFreeHost(g->h_pin);
  g->h_pin = nullptr;
  g->h_pin_bytes = 0;
  CK(cudaMallocHost((void**)&g->h_pin, bytes));
  g->h_pin_bytes = bytes;
  return TLC_OK;
}

static int ensure_call_buffers(tlc_graph* g, int64_t E) {
  if (E <= g->call_cap) return TLC_OK;
  if (g->d_n) cudaFree(g->d_n);
  if (g->d_m) cudaFree(g->d_m);
  if (g->d_ds) cudaFree(g->d_ds);
  if (g->d_st) cudaFree(g->d_st);
  if (g->d_bytes) cudaFree(g->d_bytes);
  g->d_n = g->d_m = g->d_ds = nullptr; g->d_st = nullptr; g->d_bytes = nullptr; g->call_cap = 0;
  const int64_t cap = E + E / 8 + 1024;
  CK(cudaMalloc((void**)&g->d_n, cap * 4));
  CK(cudaMalloc((void**)&g->d_m, cap * 4));
  CK(cudaMalloc((void**)&g->d_ds, cap * 4));
  CK(cudaMalloc((void**)&g->d_st, cap));
  CK(cudaMalloc((void**)&g->d_bytes, cap * 8));
  g->call_cap = cap;
  return TLC_OK;
}

static int ensure_io_buffers(tlc_graph* g, int64_t E, int r2) {
  if (E > g->io_cap_t) {
    if (g->io_t) cudaFree(g->io_t);
    if (g->io_st) cudaFree(g->io_st);
    g->io_t = nullptr; g->io_st = nullptr; g->io_cap_t = 0;
    const int64_t cap = E + E / 8 + 1024;
    CK(cudaMalloc((void**)&g->io_t, (size_t)cap * 8));
    CK(cudaMalloc((void**)&g->io_st, (size_t)cap));
    g->io_cap_t = cap;
  }
  if (E * r2 > g->io_cap_pi) {
    if (g->io_pi) cudaFree(g->io_pi);
    g->io_pi = nullptr; g->io_cap_pi = 0;
    const int64_t cap = (E + E / 8 + 1024) * r2;
    CK(cudaMalloc((void**)&g->io_pi, (size_t)cap * 8));
    g->io_cap_pi = cap;
  }
  return TLC_OK;
}

// grow-only arena: at least `need_min` (one chunk must fit), preferably `need_all` (the whole call in one chunk), never
// more than the configured size (tlc_graph_create's arena_bytes, TLC_ARENA_GB, default 24 GiB) or 85 % of free memory
static int ensure_arena(tlc_graph* g, size_t need_min, size_t need_all) {
  if (g->arena && g->arena_bytes >= need_min) {
    // large enough for a chunk; grow only if the call would otherwise be split and there is headroom left
    size_t limit = g->arena_req;
    if (limit == 0) { const char* env = getenv("TLC_ARENA_GB"); limit = (size_t)((env ? atof(env) : 24.0) * (double)(1ull << 30)); }
    if (g->arena_bytes >= std::min(need_all, limit)) return TLC_OK;
  }
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const size_t had = g->arena_bytes;
  if (g->arena) { free_b += g->arena_bytes; cudaFree(g->arena); g->arena = nullptr; g->arena_bytes = 0; }
  size_t want = g->arena_req;
  if (want == 0) {
    const char* env = getenv("TLC_ARENA_GB");
    const double gb = env ? atof(env) : 24.0;
    want = (size_t)(gb * (double)(1ull << 30));
  }
  want = std::max(std::min(want, need_all + need_all / 2), need_min);  // (no 24 GiB for a handful of small vicinities; 50 % headroom: calls of similar size do not reallocate)
  // a floor and geometric growth: the handful of large vicinities kernel S leaves to the staged pipeline differs a lot
  // from call to call, and every reallocation is a device-wide synchronisation (measured: 10 - 60 ms per call)
  {
    size_t limit = g->arena_req;
    if (limit == 0) { const char* env = getenv("TLC_ARENA_GB"); limit = (size_t)((env ? atof(env) : 24.0) * (double)(1ull << 30)); }
    want = std::max(want, std::min(limit, std::max<size_t>(2 * had, (size_t)256 << 20)));
  }
  const size_t cap = (size_t)((double)free_b * 0.85);
  if (want > cap) want = cap;
  if (want < need_min)
    return fail(TLC_E_NOMEM, "one vicinity needs " + std::to_string(need_min) + " arena bytes, only " +
                                 std::to_string(cap) + " available");
  CK(cudaMalloc((void**)&g->arena, want));
  g->arena_bytes = want;
  return TLC_OK;
}

static VicinityScratch make_vs(const tlc_graph* g) {
  return VicinityScratch{g->bitmaps, g->queue, g->vic_grid, g->ball_cache, g->ball_acc, g->ball_state, g->ball_list,
                         g->work_counter + 2};
}

static int ensure_vicinity_scratch(tlc_graph* g, const Params& p) {
  if (g->vic_hop == p.hop && g->vic_grid > 0) return TLC_OK;
  if (g->bitmaps) { cudaFree(g->bitmaps); g->bitmaps = nullptr; }
  if (g->queue) { cudaFree(g->queue); g->queue = nullptr; }
  if (g->ball_cache) { cudaFree(g->ball_cache); g->ball_cache = nullptr; }
  if (g->ball_acc) { cudaFree(g->ball_acc); g->ball_acc = nullptr; }
  if (g->ball_state) { cudaFree(g->ball_state); g->ball_state = nullptr; }
  if (g->ball_list) { cudaFree(g->ball_list); g->ball_list = nullptr; }
  {
    // one bitmap row per node, if that is affordable (TLC_BALL_CACHE_GB, default 16; 0 disables)
    const char* env = getenv("TLC_BALL_CACHE_GB");
    const double gb = env ? atof(env) : 16.0;
    const size_t Wc = ((size_t)g->gv.N + 31) / 32;
    const size_t need = (size_t)g->gv.N * Wc * sizeof(uint32_t);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (gb > 0 && (double)need <= gb * (double)(1ull << 30) && need <= free_b / 4) {
      CK(cudaMalloc((void**)&g->ball_cache, std::max<size_t>(need, 16)));
      CK(cudaMalloc((void**)&g->ball_acc, (size_t)g->gv.N * 2 * sizeof(unsigned long long)));
      CK(cudaMalloc((void**)&g->ball_state, (size_t)g->gv.N * sizeof(int32_t)));
      CK(cudaMalloc((void**)&g->ball_list, (size_t)g->gv.N * sizeof(int32_t)));
      CK(cudaMemset(g->ball_state, 0, (size_t)g->gv.N * sizeof(int32_t)));
    }
  }
  size_t W = 0;
  bool smem = false;
  g->vic_grid = vicinity_grid(g->device, g->gv, p, &W, &smem);
  if (!smem) CK(cudaMalloc((void**)&g->bitmaps, (size_t)g->vic_grid * 2 * W * sizeof(uint32_t)));
  if (p.hop > 2) CK(cudaMalloc((void**)&g->queue, (size_t)g->vic_grid * 2 * (size_t)g->gv.N * sizeof(int32_t)));
  g->vic_hop = p.hop;
  return TLC_OK;
}

// kernel 1t's tables: four [N][N] arrays, allocated once when the graph is small enough (N <= 16384: the build kernel keeps a
// root's whole distance / parent vector in shared memory) and 28 N^2 bytes fit TLC_SSSP_CACHE_GB (default 8, 0 disables)
static int ensure_sssp_tables(tlc_graph* g, const Params& p) {
  const int plain = (p.flags & TLC_F_SUM_PLAIN) ? 1 : 0;
  if (!g->sssp_tried) {
    g->sssp_tried = true;
    const char* env = getenv("TLC_SSSP_CACHE_GB");
    const double gb = env ? atof(env) : 8.0;
    const size_t N = (size_t)g->gv.N;
    const size_t need = N * N * 28;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (gb > 0 && N <= 16384 && (double)need <= gb * (double)(1ull << 30) && need <= free_b / 3 &&
        sssp_build_smem((int)N) <= 227 * 1024) {
      CK(cudaMalloc((void**)&g->sssp.D, N * N * 8));
      CK(cudaMalloc((void**)&g->sssp.Q, N * N * 8));
      CK(cudaMalloc((void**)&g->sssp.P, N * N * 4));
      CK(cudaMalloc((void**)&g->sssp.PW, N * N * 8));
      CK(cudaMalloc((void**)&g->sssp.state, N * 4));
      CK(cudaMalloc((void**)&g->sssp.list, N * 4));
      CK(cudaMalloc((void**)&g->sssp.count, 64));
      CK(cudaMalloc((void**)&g->sssp_pw, (size_t)g->sm_count * N * 8));
      CK(cudaMemset(g->sssp.state, 0, N * 4));
      g->sssp_plain = plain;
    }
  }
  if (g->sssp.D && g->sssp_plain != plain) {  // the path sums depend on the summation rule: rebuild
    CK(cudaMemsetAsync(g->sssp.state, 0, (size_t)g->gv.N * 4, g->stream));
    g->sssp_plain = plain;
    g->sssp_all = false;
  }
  return TLC_OK;
}

// carve the arena for a chunk with T targets, Nv vertices, Ne edges (pairs = Nv + Ne + T)
static size_t chunk_bytes(int64_t T, int64_t Nv, int64_t Ne, int64_t Na, int64_t Wd = 0) {
  const int64_t Np = Nv + Ne + T;
  // every array individually aligned: 16 vertex arrays, 13 edge arrays, 5 pair arrays, 16 target arrays;
  // Wd > 0 (graph-row route): bitmap + word prefix of every target
  return (size_t)(Nv * BV + Ne * BE + Na * BA + Np * BP + (T + 1) * BT + T * Wd * 8) + ALIGN * 60;
}

template <typename T_>
static T_* take(char*& cur, int64_t count) {
  T_* p = reinterpret_cast<T_*>(cur);
  cur += align_up((size_t)count * sizeof(T_));
  return p;
}

static ChunkView carve(char* arena, int64_t T, int64_t Nv, int64_t Ne, int64_t Na, const int32_t* d_targets,
                       int64_t Wd = 0) {
  ChunkView c{};
  char* cur = arena;
  const int64_t Np = Nv + Ne + T;
  c.T = (int32_t)T;
  c.tgt = d_targets;
  c.tidx = take<int64_t>(cur, T);
  c.voff = take<int64_t>(cur, T + 1);
  c.eoff = take<int64_t>(cur, T + 1);
  c.aoff = take<int64_t>(cur, T + 1);
  c.tn = take<int32_t>(cur, T); c.tm = take<int32_t>(cur, T); c.tlu = take<int32_t>(cur, T); c.tlv = take<int32_t>(cur, T);
  c.tnp = take<int32_t>(cur, T); c.tnpos = take<int32_t>(cur, T); c.tnneg = take<int32_t>(cur, T); c.tncls = take<int32_t>(cur, T);
  c.tnb = take<int32_t>(cur, T); c.tminv = take<int32_t>(cur, T); c.tmaxv = take<int32_t>(cur, T);
  c.tstatus = take<uint8_t>(cur, T);
  c.tfb = take<uint8_t>(cur, T);
  c.vord = take<int32_t>(cur, Nv); c.vrank = take<int32_t>(cur, Nv);
  c.bfirst = take<int32_t>(cur, Nv + T);
  c.astart = take<int32_t>(cur, Nv); c.adeg = take<int32_t>(cur, Nv); c.aminw = take<float>(cur, Nv);
  c.anb = take<uint32_t>(cur, Na); c.aw = take<double>(cur, Na);
  c.vert = take<int32_t>(cur, Nv); c.vcls = take<int32_t>(cur, Nv); c.vs0 = take<int32_t>(cur, Nv);
  c.vs1 = take<int32_t>(cur, Nv); c.vs2 = take<int32_t>(cur, Nv); c.neg = take<int32_t>(cur, Nv);
  c.fval = take<double>(cur, Nv); c.d1 = take<double>(cur, Nv); c.d2 = take<double>(cur, Nv);
  c.v64a = take<unsigned long long>(cur, Nv); c.v64b = take<unsigned long long>(cur, Nv);
  c.v64c = take<unsigned long long>(cur, Nv);
  c.elo = take<int32_t>(cur, Ne); c.ehi = take<int32_t>(cur, Ne); c.pos = take<int32_t>(cur, Ne);
  c.arank = take<int32_t>(cur, Ne);
  c.ew = take<double>(cur, Ne);
  c.ord_asc = take<uint32_t>(cur, Ne); c.ord_desc = take<uint32_t>(cur, Ne);
  c.sp0 = take<uint32_t>(cur, Ne); c.sp1 = take<uint32_t>(cur, Ne);
  c.sk0 = take<unsigned long long>(cur, Ne); c.sk1 = take<unsigned long long>(cur, Ne);
  c.isneg = take<uint8_t>(cur, Ne);
  c.pkind = take<uint8_t>(cur, Np);
  c.pbv = take<int32_t>(cur, Np); c.pdv = take<int32_t>(cur, Np);
  c.pbirth = take<double>(cur, Np); c.pdeath = take<double>(cur, Np);
  c.dbm = Wd > 0 ? take<uint32_t>(cur, T * 2 * Wd) : nullptr;
  c.W = (int32_t)Wd;
  return c;
}

static int block_for(int64_t m_max) {
  if (m_max >= 32768) return 512;
  if (m_max >= 4096) return 256;
  if (m_max >= 512) return 128;
  if (m_max >= 96) return 64;
  return 32;
}
static int size_class(int64_t m) { return block_for(m); }
// kernel 1b keeps 9 - 11 bytes of state per vertex in shared memory: 0 = does not fit at all (state in the arena),
// 1 = fits only in the lean layout, 2 = fits (thresholds a little under the kernel's own limits, which also hold bitmaps)
static int smem_class(int64_t n) { return n > 22000 ? 0 : (n > 16000 ? 1 : 2); }

struct StageTimer {
  // per-stage CUDA events (TLC_STAGE_TIMING=1).  The events come from a pool owned by the graph and are reused from call
  // to call: no cudaEventCreate / cudaEventDestroy inside a timed call once the pool has grown to a call's needs.
  bool on;
  cudaStream_t st;
  std::vector<cudaEvent_t>* pool;
  size_t used, first;  // next free pool slot; first slot of the marks not collected yet
  std::vector<int> stage;
  StageTimer(bool on_, cudaStream_t s, std::vector<cudaEvent_t>* pool_, size_t start) : on(on_), st(s), pool(pool_), used(start), first(start) {}
  cudaEvent_t take() {
    if (used == pool->size()) { cudaEvent_t e; cudaEventCreate(&e); pool->push_back(e); }
    return (*pool)[used++];
  }
  void mark(int stage_id) {  // marks the START of stage_id (or the end marker with id -1)
    if (!on) return;
    cudaEventRecord(take(), st);
    stage.push_back(stage_id);
  }
  void collect(double* ms8) {  // ms8: per-stage accumulators (10 slots); call after the stream is synchronised
    if (!on) return;
    for (size_t i = 0; i + 1 < stage.size(); i++) {
      if (stage[i] < 0) continue;
      float ms = 0;
      cudaEventElapsedTime(&ms, (*pool)[first + i], (*pool)[first + i + 1]);
      ms8[stage[i]] += ms;
    }
    first = used;
    stage.clear();
  }
};

// run the stage kernels over one carved chunk (offsets already uploaded)
// kernels whose shared-memory footprint follows the vicinity size are launched per SUB-RANGE of the chunk's
// (size-sorted) targets, each with the capacity of its own largest vicinity: smaller vicinities get more
// resident CTAs per SM
struct SubRange { int t0, cnt; int64_t n_max; };

static void run_stages(tlc_graph* g, const Params& p, const ChunkView& c, int64_t n_max, int64_t m_max,
                       const std::vector<SubRange>& subs, double* d_pi,
                       float* d_pi32, uint8_t* d_status, bool want_lists, StageTimer& tm, bool use_table = false) {
  cudaStream_t st = g->stream;
  const int block = block_for(m_max);
  VicinityScratch vs = make_vs(g);
  tm.mark(1);
  if (!c.dbm) launch_vicinity_fill(g->gv, p, c, vs, g->work_counter, st);
  tm.mark(2);
  if (p.flags & TLC_F_FILT_HKS) launch_hks_filtration(p, c, g->hks_time, n_max, st);
  else if (p.flags & (TLC_F_FILT_DEGREE | TLC_F_FILT_CENTRALITY | TLC_F_FILT_CLUSTERING)) launch_degree_filtration(p, c, st);
  else {
    const bool fork = subs.size() > 1 && g->ev_fork != nullptr;
    if (fork) cudaEventRecord(g->ev_fork, st);
    for (size_t i = 0; i < subs.size(); i++) {
      const SubRange& r = subs[i];
      cudaStream_t s = st;
      if (fork && i > 0 && i <= 3) { s = g->side[i - 1]; cudaStreamWaitEvent(s, g->ev_fork, 0); }
      if (c.dbm && use_table && filtration_table_smem(g->gv.N, r.n_max) <= 226 * 1024) {
        launch_filtration_table(g->gv, p, c, vs, g->sssp, r.t0, r.cnt, block, r.n_max, s);
        g->last_table += r.cnt;
      } else if (c.dbm) launch_filtration_direct(g->gv, p, c, vs, g->gminw, r.t0, r.cnt, block, r.n_max, s);
      else launch_filtration(p, c, r.t0, r.cnt, block, r.n_max, s);
      if (s != st) { cudaEventRecord(g->ev_join[i - 1], s); cudaStreamWaitEvent(st, g->ev_join[i - 1], 0); }
    }
  }
  // ascending sweep: vertex-ordered kernels 2v + 3v; targets they hand back (tfb) and, when the descending
  // sweep is wanted (diagram output, Pos/Neg lists for the loops), the edge-sorted kernels 2 + 3.  The image
  // only needs PD_up and [min,max]: PD_down and [max,min] have death <= birth, i.e. weight 0.
  const bool ext = (p.flags & TLC_F_EXTENDED) != 0;
  const bool want_desc = (ext || want_lists) && !(p.flags & TLC_F_ASC_ONLY);
  // keep-zero calls (PDGNN generators) that sort the edges anyway for the descending sweep: the vertex-ordered sweep has
  // no trivial blocks there (every merge emits a pair), so the ascending sweep rides on the edge-sorted kernels too
  const bool edge_sorted = (p.flags & TLC_F_EDGE_SORTED) != 0 || ((p.flags & TLC_F_KEEP_ZERO) != 0 && want_desc);
  tm.mark(3);
  if (edge_sorted) {
    cudaMemsetAsync(c.tfb, 1, (size_t)c.T, st);
  } else {
    {  // per sub-range like kernel 1b: the block table in shared memory is sized by the sub-range's largest vicinity
      const bool fork = subs.size() > 1 && g->ev_fork != nullptr;
      if (fork) cudaEventRecord(g->ev_fork, st);
      for (size_t i = 0; i < subs.size(); i++) {
        const SubRange& r = subs[i];
        cudaStream_t s = st;
        if (fork && i > 0 && i <= 3) { s = g->side[i - 1]; cudaStreamWaitEvent(s, g->ev_fork, 0); }
        launch_vorder(p, c, r.t0, r.cnt, block, r.n_max, s);
        if (s != st) { cudaEventRecord(g->ev_join[i - 1], s); cudaStreamWaitEvent(st, g->ev_join[i - 1], 0); }
      }
    }
    tm.mark(4);
    {  // likewise kernel 3v: its shared-memory parents are sized by the sub-range's largest vicinity
      const bool fork = subs.size() > 1 && g->ev_fork != nullptr;
      if (fork) cudaEventRecord(g->ev_fork, st);
      for (size_t i = 0; i < subs.size(); i++) {
        const SubRange& r = subs[i];
        cudaStream_t s = st;
        if (fork && i > 0 && i <= 3) { s = g->side[i - 1]; cudaStreamWaitEvent(s, g->ev_fork, 0); }
        launch_sweep(p, c, r.t0, r.cnt, r.n_max, s);
        if (s != st) { cudaEventRecord(g->ev_join[i - 1], s); cudaStreamWaitEvent(st, g->ev_join[i - 1], 0); }
      }
    }
  }
  tm.mark(5);
  if (!c.dbm) {  // (graph-row route: no adjacency to derive edges from; a chunk with handed-back targets is redone)
    // the edge-sorted kernels work on the canonical (lo, hi) edge list, derived from the adjacency where needed
    launch_edgelist(p, c, block, want_desc ? 0 : 1, st);
    // kernel 3b ranks loop edges by their position in the ascending order, so `extended` needs ord_asc of every target
    launch_sort(p, c, block, want_desc ? 3 : 1, want_desc ? 0 : 1, st);
  }
  tm.mark(6);
  const int64_t uf_ints = 2 * n_max * 4 <= 96 * 1024 ? 2 * n_max : 0;
  if (!c.dbm) launch_union_find(p, c, block, (int)uf_ints, want_desc ? 1 : 0, want_desc ? 3 : 1, 1, st);
  tm.mark(7);
  if (ext) {
    const int64_t lp_ints = 3 * n_max * 4 <= 200 * 1024 ? 3 * n_max : 0;
    launch_loops(p, c, std::min(block, 128), (int)lp_ints, n_max, st);
  }
  tm.mark(8);
  launch_pimg(p, c, d_pi, d_pi32, d_status, std::max(32, std::min(block, 256)), st);
  tm.mark(-1);
}

static int check_params(const tlc_params* p) {
  if (!p) return fail(TLC_E_INVALID, "params is NULL");
  if (p->resolution < 1 || p->resolution > 16) return fail(TLC_E_INVALID, "resolution must be in 1..16");
  if (p->hop < 0) return fail(TLC_E_INVALID, "hop must be >= 0");
  if (p->mode < TLC_MODE_EDGE || p->mode > TLC_MODE_EDGE_REMOVEINTER) return fail(TLC_E_INVALID, "bad mode");
  if ((p->flags & TLC_F_ASC_ONLY) && (p->flags & TLC_F_EXTENDED))
    return fail(TLC_E_INVALID, "TLC_F_ASC_ONLY excludes TLC_F_EXTENDED (the loops need the Pos/Neg lists)");
  return TLC_OK;
}

// the STAGED pipeline over device-resident targets.  detail != NULL: single chunk in input order, every
// intermediate copied back to host.
static int run_staged(tlc_graph* g, const int32_t* d_targets, int64_t E, const tlc_params* up, double* d_pi,
                      float* d_pi32, uint8_t* d_status, int64_t* cnt_compute, tlc_detail* detail) {
  int rc = check_params(up);
  if (rc) return rc;
  if (E < 0 || E >= ((int64_t)1 << 32)) return fail(TLC_E_INVALID, "E must be in [0, 2^32)");
  CK(cudaSetDevice(g->device));
  Params p{up->hop, up->mode, up->descriptor, up->resolution, up->flags, up->img_mask};
  const int r2 = p.resolution * p.resolution;
  cudaStream_t st = g->stream;
  for (int i = 0; i < 10; i++) g->stage_ms[i] = 0;
  g->nchunks = 0;
  g->alg_bytes = g->alg_bytes_bfs = g->alg_bytes_uf = 0;
  g->last_live = g->last_nv = g->last_ne = g->last_fb = g->last_direct = 0;
  CK(cudaMemsetAsync(g->work_counter + 4, 0, 4 * sizeof(int), st));
  if (cnt_compute) *cnt_compute = 0;
  if (E == 0) return TLC_OK;
  const bool bad_desc = p.descriptor < 0 || p.descriptor > 2;

  if ((rc = ensure_call_buffers(g, E))) return rc;
  if ((rc = ensure_vicinity_scratch(g, p))) return rc;
  const char* tenv = getenv("TLC_STAGE_TIMING");
  g->timing = tenv && atoi(tenv) != 0;
  StageTimer tm(g->timing, st, &g->ev_pool, g->ev_base);
  cudaEvent_t ev_total0 = nullptr, ev_total1 = nullptr;
  if (g->timing) { ev_total0 = tm.take(); ev_total1 = tm.take(); tm.first = tm.used; cudaEventRecord(ev_total0, st); }

  // ---- route: graph-row (no adjacency in HBM) or materialised, per call ----
  // the graph-row route serves calls that need the ascending sweep only; it reads D_S graph-row entries per root
  // where the materialised route reads the 2m induced ones, so it is taken where the vicinities are dense in the graph
  const int64_t Wd = ((int64_t)g->gv.N + 31) / 32;
  const bool want_desc_call = ((p.flags & TLC_F_EXTENDED) != 0 || detail != nullptr) && !(p.flags & TLC_F_ASC_ONLY);
  const bool direct_ok = g->ball_cache != nullptr && g->gminw != nullptr && !want_desc_call && !bad_desc && p.mode <= TLC_MODE_NODE &&
                         !(p.flags & TLC_F_FILT_ANY_STRUCT) &&
                         !(p.flags & (TLC_F_NO_DIRECT | TLC_F_EDGE_SORTED)) && (size_t)Wd * 8 <= 64 * 1024;
  double direct_ratio = 2.0;
  if (const char* env = getenv("TLC_DIRECT_RATIO")) direct_ratio = atof(env);
  // ---- kernel 1, counting pass ----
  VicinityScratch vs = make_vs(g);
  tm.mark(0);
  launch_ball_cache(g->gv, p, d_targets, E, vs, st);
  bool call_direct = direct_ok && (p.flags & TLC_F_DIRECT);
  if (direct_ok && !call_direct) {
    // density probe: the full counting pass on a strided sample of the call's targets (measured on B200: one route
    // per call beats mixing; Computers-shaped 2-hop: D_S / 2m = 1.4, PubMed / collab: > 2).  The verdict is kept
    // for the next calls with the same hop and mode and refreshed every 32 calls.
    if (g->probe_hop == p.hop && g->probe_mode == p.mode && g->probe_age < 32) {
      call_direct = g->probe_direct;
      g->probe_age++;
    } else {
      const int64_t S = std::min<int64_t>(E, std::max<int64_t>(64, E / 64));
      const int64_t stride = E / S;
      DevBuf sample_buf;
      CK(sample_buf.alloc((size_t)S * 8));
      int32_t* d_sample = sample_buf.as<int32_t>();
      CK(cudaMemcpy2DAsync(d_sample, 8, d_targets, (size_t)stride * 8, 8, (size_t)S, cudaMemcpyDeviceToDevice, st));
      launch_vicinity_sizes(g->gv, p, d_sample, S, g->d_n, g->d_m, g->d_ds, g->d_st, g->d_bytes, vs, g->work_counter, st);
      std::vector<int32_t> sm((size_t)S), sds((size_t)S);
      std::vector<uint8_t> sst((size_t)S);
      CK(cudaMemcpyAsync(sm.data(), g->d_m, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(sds.data(), g->d_ds, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(sst.data(), g->d_st, (size_t)S, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      double a = 0, b = 0;
      for (int64_t i = 0; i < S; i++) if (sst[(size_t)i] == TLC_ST_OK) { a += sds[(size_t)i]; b += 2.0 * sm[(size_t)i]; }
      call_direct = b > 0 && a <= direct_ratio * b;
      g->probe_hop = p.hop; g->probe_mode = p.mode; g->probe_age = 0; g->probe_direct = call_direct;
    }
  }
  // the batch call on the graph-row route skips counting the induced edges (kernel 1b counts them as it reads the rows)
  const bool light = call_direct && detail == nullptr;
  // kernel 1t: the graph-row route's filtration from per-root tables of the whole graph, where they are affordable
  bool use_table = false;
  g->last_table = 0;
  if (call_direct && !(p.flags & TLC_F_NO_TABLE) && !getenv("TLC_NO_TABLE")) {
    if ((rc = ensure_sssp_tables(g, p))) return rc;
    if (g->sssp.D) {
      use_table = true;
      if (!g->sssp_all) {
        // the rows of ALL roots, once per graph (and summation rule): a full pass over a dataset touches every node anyway,
        // and a handful of late rows would cost a whole single-CTA shortest-path run in the middle of a step
        cudaEvent_t b0 = nullptr, b1 = nullptr;
        cudaEventCreate(&b0); cudaEventCreate(&b1);
        cudaEventRecord(b0, st);
        launch_sssp_build(g->gv, p, nullptr, 0, g->sssp, g->gminw, g->sssp_pw, g->sm_count, st);
        cudaEventRecord(b1, st);
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, b0, b1);
        g->sssp_build_ms = ms;
        cudaEventDestroy(b0); cudaEventDestroy(b1);
        g->sssp_all = true;
      }
    }
  }
  if (light)
    launch_vicinity_light(g->gv, p, d_targets, E, g->d_n, g->d_m, g->d_ds, g->d_st, g->d_bytes, vs, g->sm_count, st);
  else
    launch_vicinity_sizes(g->gv, p, d_targets, E, g->d_n, g->d_m, g->d_ds, g->d_st, g->d_bytes, vs, g->work_counter, st);
  tm.mark(-1);
  const size_t pin_need = align_up((size_t)E * 8) + (size_t)E * 9 + ALIGN;
  const size_t pin_need2 = align_up((size_t)E * 4);  // h_ds
  if ((rc = ensure_pinned(g, align_up(pin_need) + pin_need2 + (size_t)(E + 2) * 8 * 4 + (size_t)(E + 2) * 4 + ALIGN))) return rc;
  int32_t* h_n = reinterpret_cast<int32_t*>(g->h_pin);
  int32_t* h_m = h_n + E;
  double* h_bytes = reinterpret_cast<double*>(g->h_pin + align_up((size_t)E * 8));
  uint8_t* h_st = reinterpret_cast<uint8_t*>(h_bytes + E);
  int32_t* h_ds = reinterpret_cast<int32_t*>(g->h_pin + align_up(pin_need));
  char* h_chunk = g->h_pin + align_up(pin_need) + pin_need2;  // [tidx | voff | eoff | aoff] staging, (E+2)*8 each
  CK(cudaMemcpyAsync(h_n, g->d_n, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_m, g->d_m, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_ds, g->d_ds, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_st, g->d_st, (size_t)E, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_bytes, g->d_bytes, (size_t)E * 8, cudaMemcpyDeviceToHost, st));
  // rows of targets that never reach the image kernel stay zero (riccidist2dgm.py:363)
  CK(cudaMemsetAsync(d_pi, 0, (size_t)E * r2 * sizeof(double), st));
  if (d_pi32) CK(cudaMemsetAsync(d_pi32, 0, (size_t)E * r2 * sizeof(float), st));
  CK(cudaStreamSynchronize(st));
  if (bad_desc) {  // KeyError in perturb_filter_function for every target that got that far (accelerated_PD.py:13)
    for (int64_t i = 0; i < E; i++) if (h_st[i] == TLC_ST_OK) h_st[i] = TLC_ST_BAD_DESCRIPTOR;
    if (d_status) CK(cudaMemcpyAsync(d_status, h_st, (size_t)E, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return TLC_OK;
  }
  if (d_status) CK(cudaMemcpyAsync(d_status, g->d_st, (size_t)E, cudaMemcpyDeviceToDevice, st));
  for (int64_t i = 0; i < E; i++)
    if (h_st[i] == TLC_ST_OK) { g->alg_bytes += h_bytes[i]; g->last_live++; g->last_nv += h_n[i]; if (!light) g->last_ne += h_m[i]; }

  auto is_direct = [&](int64_t) { return call_direct; };
  // ---- plan chunks ----
  std::vector<int64_t> order;
  order.reserve(E);
  bool detail_direct = false;
  if (detail) {
    for (int64_t i = 0; i < E; i++) order.push_back(i);
    detail_direct = direct_ok && (p.flags & TLC_F_DIRECT);  // one chunk, one route
  } else {
    // one 64-bit key per live target, then a plain sort: [route | shared-memory class | size, largest first | row]
    // (a comparator-based stable sort of the index vector cost ~1 ms per 8 k targets: a third of a PubMed-shaped step).
    // Graph-row targets first (chunks are route-homogeneous); ahead of everything the few vicinities too large for
    // kernel 1b's shared-memory state (heavy-tailed graphs): they get sub-ranges of their own (below) instead of
    // dragging a whole size group onto the arena path.
    std::vector<uint64_t> keys;
    keys.reserve(E);
    for (int64_t i = 0; i < E; i++) {
      if (h_st[i] != TLC_ST_OK) continue;
      const uint64_t route = is_direct(i) ? 0 : 1, cls = (uint64_t)smem_class(h_n[i]);
      const uint64_t msz = (uint64_t)std::min<int64_t>(h_m[i], (1 << 28) - 1);
      keys.push_back((route << 63) | (cls << 61) | ((((uint64_t)1 << 28) - 1 - msz) << 32) | (uint64_t)i);
    }
    std::sort(keys.begin(), keys.end());
    for (uint64_t k : keys) order.push_back((int64_t)(k & 0xffffffffull));
  }
  int64_t need_one = 0;
  for (int64_t i : order)
    need_one = std::max<int64_t>(need_one, (int64_t)chunk_bytes(1, h_n[i], light ? 0 : h_m[i], light ? 0 : h_ds[i],
                                                                is_direct(i) ? Wd : 0));
  if (detail) {
    int64_t Nv = 0, Ne = 0, Na = 0;
    for (int64_t i : order) { Nv += h_n[i]; Ne += h_m[i]; Na += h_ds[i]; }
    need_one = (int64_t)chunk_bytes(E, Nv, Ne, Na, detail_direct ? Wd : 0);
    if (Nv > detail->cap_v || Ne > detail->cap_e || Nv + Ne + E > detail->cap_p)
      return fail(TLC_E_CAPACITY, "detail capacity too small: need v=" + std::to_string(Nv) + " e=" +
                                      std::to_string(Ne) + " p=" + std::to_string(Nv + Ne + E));
  }
  size_t need_all = (size_t)need_one;
  if (!detail) {
    int64_t Nv = 0, Ne = 0, Na = 0;
    for (int64_t i : order) { Nv += h_n[i]; if (!light) { Ne += h_m[i]; Na += h_ds[i]; } }
    need_all = std::max(need_all, chunk_bytes((int64_t)order.size(), Nv, Ne, Na, call_direct ? Wd : 0));
  }
  if ((rc = ensure_arena(g, (size_t)need_one, need_all))) return rc;

  const int64_t max_T = 1 << 16;
  std::vector<int64_t> redo;  // graph-row batch call: rows of the targets kernel 3v handed back
  size_t pos = 0;
  int64_t* h_tidx = reinterpret_cast<int64_t*>(h_chunk);
  int64_t* h_voff = h_tidx + (E + 2);
  int64_t* h_eoff = h_voff + (E + 2);
  int64_t* h_aoff = h_eoff + (E + 2);
  int32_t* h_tm = reinterpret_cast<int32_t*>(h_aoff + (E + 2));
  while (pos < order.size()) {
    // greedy pack: same route, same size class, fits the arena
    int64_t T = 0, Nv = 0, Ne = 0, Na = 0, n_max = 0, m_max = 0;
    int cls0 = size_class(h_m[order[pos]]);
    const bool direct = detail ? detail_direct : is_direct(order[pos]);
    const int64_t Wc = direct ? Wd : 0;
    size_t q = pos;
    while (q < order.size() && T < max_T) {
      const int64_t i = order[q];
      if (!detail && is_direct(i) != direct) break;
      if (!detail && size_class(h_m[i]) != cls0) {
        // a new size class starts its own chunk (own CTA width) only if it can fill the GPU about twice;
        // a handful of smaller targets ride along in the current chunk instead of paying a tail of their own
        const int cls1 = size_class(h_m[i]);
        size_t q2 = q;
        while (q2 < order.size() && q2 - q < (size_t)(2 * g->sm_count) && size_class(h_m[order[q2]]) == cls1 &&
               is_direct(order[q2]) == direct) q2++;
        if (q2 - q >= (size_t)(2 * g->sm_count)) break;
        cls0 = cls1;
      }
      const int64_t mi = light ? 0 : h_m[i], di = light ? 0 : h_ds[i];  // (graph-row batch call: no edge / adjacency arrays)
      if (chunk_bytes(T + 1, Nv + h_n[i], Ne + mi, Na + di, Wc) > g->arena_bytes) break;
      // the staging buffers of a chunk are reused: wait for the previous chunk's upload (stream order suffices,
      // the pinned region of this chunk is [pos, q) which no earlier chunk touches)
      h_tidx[q] = i; h_voff[q] = Nv; h_eoff[q] = Ne; h_aoff[q] = Na; h_tm[q] = h_m[i];
      Nv += h_n[i]; Ne += mi; Na += di;
      n_max = std::max<int64_t>(n_max, h_n[i]); m_max = std::max<int64_t>(m_max, h_m[i]);
      T++; q++;
    }
    if (T == 0) return fail(TLC_E_NOMEM, "a vicinity does not fit the arena");
    ChunkView c = carve(g->arena, T, Nv, Ne, Na, d_targets, Wc);
    c.fb_counter = g->work_counter + 4;
    c.gcol = g->gv.col;
    c.gkappa = g->gv.kappa;
    // offsets: [pos, pos+T) plus the terminating total.  voff/eoff need T+1 entries; the terminator is
    // written into a separate tiny pinned slot so that the next chunk's slot `q` is not clobbered.
    CK(cudaMemcpyAsync((void*)c.tidx, h_tidx + pos, (size_t)T * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)c.voff, h_voff + pos, (size_t)T * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)c.eoff, h_eoff + pos, (size_t)T * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)c.aoff, h_aoff + pos, (size_t)T * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)(c.voff + T), &Nv, 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)(c.eoff + T), &Ne, 8, cudaMemcpyHostToDevice, st));
    std::vector<SubRange> subs;
    {
      int groups = T >= 4 * 2 * g->sm_count ? 4 : (T >= 2 * 2 * g->sm_count ? 2 : 1);
      if (const char* env = getenv("TLC_SUBGROUPS")) groups = std::max(1, std::min(4, atoi(env)));  // (tuning experiments)
      // the shared-memory classes first (a chunk is sorted by them), then the rest in `groups` equal parts
      int64_t first = 0;
      for (int cls = 0; cls < 2; cls++) {
        int64_t b = first;
        while (b < T && smem_class(h_n[order[pos + b]]) == cls) b++;
        if (b > first) {
          int64_t nm = 0;
          for (int64_t k = first; k < b; k++) nm = std::max<int64_t>(nm, h_n[order[pos + k]]);
          subs.push_back(SubRange{(int)first, (int)(b - first), nm});
          first = b;
        }
      }
      const int64_t rest = T - first;
      for (int gi = 0; gi < groups && rest > 0; gi++) {
        const int64_t a = first + rest * gi / groups, b = first + rest * (gi + 1) / groups;
        if (b <= a) continue;
        int64_t nm = 0;
        for (int64_t k = a; k < b; k++) nm = std::max<int64_t>(nm, h_n[order[pos + k]]);
        subs.push_back(SubRange{(int)a, (int)(b - a), nm});
      }
    }
    if (direct) {
      c.count_m = light ? 1 : 0;
      // tm: counted by kernel 1b (batch call) or known from the full counting pass (detail call)
      if (!light) CK(cudaMemcpyAsync((void*)c.tm, h_tm + pos, (size_t)T * 4, cudaMemcpyHostToDevice, st));
      int fb0 = 0, fb1 = 0;
      CK(cudaMemcpyAsync(&fb0, g->work_counter + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
      run_stages(g, p, c, n_max, m_max, subs, d_pi, d_pi32, d_status, detail != nullptr, tm, use_table);
      CK(cudaMemcpyAsync(&fb1, g->work_counter + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
      if (light) CK(cudaMemcpyAsync(h_tm + pos, c.tm, (size_t)T * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      g->last_direct += T;
      if (light) {  // the exact edge counts: totals and the 16 m term of the compulsory bytes
        for (int64_t k = 0; k < T; k++) { g->last_ne += h_tm[pos + k]; g->alg_bytes += 16.0 * (double)h_tm[pos + k]; }
      }
      if (fb1 != fb0) {
        // kernel 3v handed targets back: their edge-sorted sweep needs the adjacency
        if (light) {  // no adjacency space was reserved: those targets are redone on the materialised route below
          std::vector<uint8_t> fbv((size_t)T);
          CK(cudaMemcpyAsync(fbv.data(), c.tfb, (size_t)T, cudaMemcpyDeviceToHost, st));
          CK(cudaStreamSynchronize(st));
          for (int64_t k = 0; k < T; k++) if (fbv[(size_t)k]) redo.push_back(h_tidx[pos + k]);
        } else {
        // the chunk is redone on the materialised route (its space is reserved), image rows and statuses are rewritten
        c.dbm = nullptr;
        c.W = 0;
        g->last_direct -= T;
        run_stages(g, p, c, n_max, m_max, subs, d_pi, d_pi32, d_status, detail != nullptr, tm);
        }
      }
    } else {
      run_stages(g, p, c, n_max, m_max, subs, d_pi, d_pi32, d_status, detail != nullptr, tm);
    }
    g->nchunks++;
    if (detail) {
      // copy every intermediate back (input order == chunk order here)
      CK(cudaStreamSynchronize(st));
      g->detail_chunk = c;
#define D2H(dst, src, count, type) \
  if (detail->dst) CK(cudaMemcpy(detail->dst, src, (size_t)(count) * sizeof(type), cudaMemcpyDeviceToHost))
      D2H(n, c.tn, T, int32_t); D2H(m, c.tm, T, int32_t); D2H(lu, c.tlu, T, int32_t); D2H(lv, c.tlv, T, int32_t);
      D2H(npairs, c.tnp, T, int32_t); D2H(npos, c.tnpos, T, int32_t); D2H(nneg, c.tnneg, T, int32_t);
      D2H(vert, c.vert, Nv, int32_t); D2H(fval, c.fval, Nv, double); D2H(neg, c.neg, Nv, int32_t);
      if (!(p.flags & TLC_F_ASC_ONLY)) {
        D2H(elo, c.elo, Ne, int32_t); D2H(ehi, c.ehi, Ne, int32_t); D2H(ew, c.ew, Ne, double);
        D2H(ord_asc, c.ord_asc, Ne, int32_t); D2H(ord_desc, c.ord_desc, Ne, int32_t); D2H(pos, c.pos, Ne, int32_t);
      }
      D2H(pbv, c.pbv, Nv + Ne + T, int32_t); D2H(pdv, c.pdv, Nv + Ne + T, int32_t);
      D2H(pbirth, c.pbirth, Nv + Ne + T, double); D2H(pdeath, c.pdeath, Nv + Ne + T, double);
#undef D2H
      if (detail->pkind) {
        std::vector<uint8_t> k8((size_t)(Nv + Ne + T));
        CK(cudaMemcpy(k8.data(), c.pkind, k8.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < k8.size(); i++) detail->pkind[i] = k8[i];
      }
      for (int64_t t = 0; t <= T; t++) {
        const int64_t vo = t < T ? h_voff[pos + t] : Nv, eo = t < T ? h_eoff[pos + t] : Ne;
        if (detail->voff) detail->voff[t] = vo;
        if (detail->eoff) detail->eoff[t] = eo;
        if (detail->poff) detail->poff[t] = vo + eo + t;
      }
    } else {
      // Nv / Ne are stack variables read by the async copies above
      CK(cudaStreamSynchronize(st));
    }
    pos = q;
  }
  if (!redo.empty()) {
    // the handed-back targets of a graph-row call, as their own call on the materialised route; rows scattered back
    const int64_t k = (int64_t)redo.size();
    // (all copies on the call's stream -- a non-blocking stream is not ordered with the legacy stream a plain cudaMemcpy uses --
    // and synchronised before the host vectors are touched or go away; the device buffers free themselves on every exit path)
    std::vector<int32_t> h_all((size_t)E * 2), h_sub((size_t)k * 2);
    CK(cudaMemcpyAsync(h_all.data(), d_targets, (size_t)E * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < k; i++) { h_sub[2 * i] = h_all[2 * redo[i]]; h_sub[2 * i + 1] = h_all[2 * redo[i] + 1]; }
    DevBuf b_sub, b_idx, b_pi, b_pi32, b_st;
    CK(b_sub.alloc((size_t)k * 8));
    CK(b_idx.alloc((size_t)k * 8));
    CK(b_pi.alloc((size_t)k * r2 * 8));
    CK(b_pi32.alloc((size_t)k * r2 * 4));
    CK(b_st.alloc((size_t)k));
    int32_t* d_sub = b_sub.as<int32_t>(); int64_t* d_idx = b_idx.as<int64_t>();
    double* s_pi = b_pi.as<double>(); float* s_pi32 = b_pi32.as<float>(); uint8_t* s_st = b_st.as<uint8_t>();
    CK(cudaMemcpyAsync(d_sub, h_sub.data(), (size_t)k * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_idx, redo.data(), (size_t)k * 8, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    // keep this call's statistics: the sub-call resets them
    const int64_t o_live = g->last_live, o_nv = g->last_nv, o_ne = g->last_ne, o_direct = g->last_direct;
    const int o_chunks = g->nchunks;
    const double o_bytes = g->alg_bytes;
    double o_ms[10];
    tm.collect(g->stage_ms);
    for (int i = 0; i < 10; i++) o_ms[i] = g->stage_ms[i];
    tlc_params up2 = *up;
    up2.flags = (up2.flags | TLC_F_NO_DIRECT) & ~TLC_F_DIRECT;
    const size_t keep_base = g->ev_base;
    g->ev_base = tm.used;  // the nested call's stage events live behind this call's
    rc = run_staged(g, d_sub, k, &up2, s_pi, s_pi32, s_st, nullptr, nullptr);
    g->ev_base = keep_base;
    if (rc == TLC_OK) {
      launch_scatter_rows(s_pi, d_pi32 ? s_pi32 : nullptr, d_status ? s_st : nullptr, d_idx, k, r2, d_pi, d_pi32, d_status,
                          g->sm_count, st);
      cudaStreamSynchronize(st);
    }
    if (rc) return rc;
    const int64_t sub_fb = g->last_fb;
    g->last_live = o_live; g->last_nv = o_nv; g->last_ne = o_ne; g->last_direct = o_direct - k; g->alg_bytes = o_bytes;
    g->nchunks += o_chunks;
    for (int i = 0; i < 10; i++) g->stage_ms[i] += o_ms[i];
    (void)sub_fb;  // (the sweep counters read back below are the sub-call's: the same k targets handed back again)
  }
  if (g->timing) {
    cudaEventRecord(ev_total1, st);
    cudaStreamSynchronize(st);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_total0, ev_total1);
    g->stage_ms[9] = ms;
    tm.collect(g->stage_ms);
  }
  // cnt_compute: targets whose image row was really computed (riccidist2dgm.py:354)
  if (cnt_compute || (detail && detail->status)) {
    if (d_status) {
      CK(cudaMemcpyAsync(h_st, d_status, (size_t)E, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      int64_t cnt = 0;
      for (int64_t i = 0; i < E; i++) cnt += h_st[i] <= TLC_ST_TRIVIAL;
      if (cnt_compute) *cnt_compute = cnt;
      if (detail && detail->status) memcpy(detail->status, h_st, (size_t)E);
    }
  }
  {
    int fb[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(fb, g->work_counter + 4, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    g->last_fb = fb[0]; g->last_general = fb[1]; g->last_rowcheck = fb[2]; g->last_blocks = fb[3];
  }
  CK(cudaGetLastError());
  return TLC_OK;
}

static int ensure_small_buffers(tlc_graph* g, int64_t E) {
  if (!g->sm_dev) {
    CK(cudaMalloc((void**)&g->sm_dev, 256));
    CK(cudaMallocHost((void**)&g->sm_host, 256));
    CK(cudaEventCreateWithFlags(&g->ev_small_a, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g->ev_small_bc, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g->ev_small_c, cudaEventDisableTiming));
    if (!getenv("TLC_NO_SIDE_STREAMS")) {
      CK(cudaStreamCreateWithFlags(&g->small_stream, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&g->small_stream2, cudaStreamNonBlocking));
    }
  }
  if (E <= g->sm_cap) return TLC_OK;
  if (g->sm_list_b) cudaFree(g->sm_list_b);
  if (g->sm_list_c) cudaFree(g->sm_list_c);
  if (g->sm_list_big) cudaFree(g->sm_list_big);
  if (g->sm_list_b2) cudaFree(g->sm_list_b2);
  g->sm_list_b = g->sm_list_c = g->sm_list_big = g->sm_list_b2 = nullptr; g->sm_cap = 0;
  const int64_t cap = E + E / 8 + 1024;
  CK(cudaMalloc((void**)&g->sm_list_b, (size_t)cap * 4));
  CK(cudaMalloc((void**)&g->sm_list_c, (size_t)cap * 4));
  CK(cudaMalloc((void**)&g->sm_list_big, (size_t)cap * 4));
  CK(cudaMalloc((void**)&g->sm_list_b2, (size_t)cap * 4));
  g->sm_cap = cap;
  return TLC_OK;
}

static int ensure_small_sub(tlc_graph* g, int64_t k, int r2) {
  if (k <= g->sm_sub_cap) return TLC_OK;
  cudaFree(g->sm_sub); cudaFree(g->sm_idx); cudaFree(g->sm_pi); cudaFree(g->sm_pi32); cudaFree(g->sm_st);
  g->sm_sub = nullptr; g->sm_idx = nullptr; g->sm_pi = nullptr; g->sm_pi32 = nullptr; g->sm_st = nullptr; g->sm_sub_cap = 0;
  const int64_t cap = k + k / 4 + 256;
  CK(cudaMalloc((void**)&g->sm_sub, (size_t)cap * 8));
  CK(cudaMalloc((void**)&g->sm_idx, (size_t)cap * 8));
  CK(cudaMalloc((void**)&g->sm_pi, (size_t)cap * r2 * 8));
  CK(cudaMalloc((void**)&g->sm_pi32, (size_t)cap * r2 * 4));
  CK(cudaMalloc((void**)&g->sm_st, (size_t)cap));
  g->sm_sub_cap = cap;
  return TLC_OK;
}

static constexpr size_t SM_STATS_OFF = 32;  // SmallStats behind the eight list counters in sm_dev / sm_host

// which calls kernel S (k0_small.cu) can take: the batch call with a Ricci-distance filtration and a 5 x 5 image on a
// graph whose ball cache exists; the route-forcing diagnostic flags keep their meaning (they name staged kernels)
static bool small_applicable(const tlc_graph* g, const tlc_params* up, int64_t E, const tlc_detail* detail) {
  if (detail || E <= 0 || E >= ((int64_t)1 << 31)) return false;
  if (up->resolution != 5 || up->descriptor < 0 || up->descriptor > 2) return false;
  if (up->flags & (TLC_F_FILT_ANY_STRUCT | TLC_F_EDGE_SORTED | TLC_F_DIRECT |
                   TLC_F_NO_DIRECT | TLC_F_NO_SMALL | TLC_F_ASC_ONLY)) return false;
  if (getenv("TLC_NO_SMALL")) return false;
  return true;
}

// statistics of a staged sub-call, accumulated across the (up to two) sub-calls of a kernel S call
struct StagedAcc {
  double ms[10] = {0};
  int64_t live = 0, nv = 0, ne = 0, fb = 0, general = 0, rowcheck = 0, blocks = 0, direct = 0, table = 0;
  int chunks = 0;
  double bytes = 0;
  void add(const tlc_graph* g) {
    for (int i = 0; i < 10; i++) ms[i] += g->stage_ms[i];
    live += g->last_live; nv += g->last_nv; ne += g->last_ne; fb += g->last_fb; general += g->last_general;
    rowcheck += g->last_rowcheck; blocks += g->last_blocks; direct += g->last_direct; table += g->last_table;
    chunks += g->nchunks; bytes += g->alg_bytes;
  }
};

// the staged pipeline over a device-side list of rows, results scattered back into the caller's tables
static int run_staged_list(tlc_graph* g, const int32_t* d_targets, const int32_t* d_list, int64_t k, const tlc_params* up,
                           double* d_pi, float* d_pi32, uint8_t* d_status, StageTimer& tm, StagedAcc& acc) {
  const int r2 = up->resolution * up->resolution;
  cudaStream_t st = g->stream;
  int rc;
  if ((rc = ensure_small_sub(g, k, r2))) return rc;
  launch_gather_targets(d_targets, d_list, k, g->sm_sub, g->sm_idx, st);
  tlc_params up2 = *up;
  up2.flags |= TLC_F_NO_SMALL;
  const size_t keep_base = g->ev_base;
  g->ev_base = tm.used;  // the nested call's stage events live behind this call's
  rc = run_staged(g, g->sm_sub, k, &up2, g->sm_pi, g->sm_pi32, g->sm_st, nullptr, nullptr);
  g->ev_base = keep_base;
  if (rc) return rc;
  launch_scatter_rows(g->sm_pi, d_pi32 ? g->sm_pi32 : nullptr, d_status ? g->sm_st : nullptr, g->sm_idx, k, r2, d_pi, d_pi32,
                      d_status, g->sm_count, st);
  CK(cudaStreamSynchronize(st));
  acc.add(g);
  return TLC_OK;
}

// kernel S over the whole call, the staged pipeline over the rows it hands on.  Class A (a warp per target) sees every
// target and sorts the rest into "classes B / C" and "more than 1024 vertices: staged"; classes B and C then run on a
// stream of their own WHILE the staged pipeline takes the big targets (whose single-lane sweeps are the critical path of
// extended calls); what class C cannot take either (<= 1024 vertices but more than 4096 edges) follows at the end.
static int run_small(tlc_graph* g, const int32_t* d_targets, int64_t E, const tlc_params* up, double* d_pi, float* d_pi32,
                     uint8_t* d_status, int64_t* cnt_compute, bool* fell_through) {
  *fell_through = false;
  CK(cudaSetDevice(g->device));
  Params p{up->hop, up->mode, up->descriptor, up->resolution, up->flags, up->img_mask};
  cudaStream_t st = g->stream;
  int rc;
  if ((rc = ensure_call_buffers(g, E))) return rc;
  if ((rc = ensure_vicinity_scratch(g, p))) return rc;
  if (!g->ball_cache) { *fell_through = true; return TLC_OK; }
  if ((rc = ensure_small_buffers(g, E))) return rc;
  if (g->small_hop != p.hop || g->small_mode != p.mode) { g->small_hop = p.hop; g->small_mode = p.mode; g->small_skip = 0; }
  if (g->small_skip > 0) { g->small_skip--; *fell_through = true; return TLC_OK; }
  const char* tenv = getenv("TLC_STAGE_TIMING");
  g->timing = tenv && atoi(tenv) != 0;
  StageTimer tm(g->timing, st, &g->ev_pool, g->ev_base);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evb0 = nullptr, ev1b = nullptr, evc0 = nullptr, ev2 = nullptr;
  if (g->timing) { ev0 = tm.take(); ev1 = tm.take(); evb0 = tm.take(); ev1b = tm.take(); evc0 = tm.take(); ev2 = tm.take(); }
  VicinityScratch vs = make_vs(g);
  int* counters = reinterpret_cast<int*>(g->sm_dev);
  SmallStats* d_stats = reinterpret_cast<SmallStats*>(g->sm_dev + SM_STATS_OFF);
  launch_ball_cache(g->gv, p, d_targets, E, vs, st);
  if (ev0) cudaEventRecord(ev0, st);
  launch_small(g->gv, p, d_targets, E, vs, d_pi, d_pi32, d_status, g->sm_list_b, g->sm_list_c, g->sm_list_big, counters,
               nullptr, nullptr, nullptr, d_stats, g->sm_count, 1, st, nullptr);
  if (ev1) cudaEventRecord(ev1, st);
  // the rows class A deferred, routed by their exact sizes: class B, class C and the staged pipeline then run side by side
  launch_small(g->gv, p, d_targets, E, vs, d_pi, d_pi32, d_status, g->sm_list_b, g->sm_list_c, g->sm_list_big, counters,
               nullptr, nullptr, nullptr, d_stats, g->sm_count, 4, st, nullptr, g->sm_list_b2);
  CK(cudaEventRecord(g->ev_small_a, st));
  CK(cudaMemcpyAsync(g->sm_host, g->sm_dev, 32, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int nB = reinterpret_cast<const int*>(g->sm_host)[4], nCc = reinterpret_cast<const int*>(g->sm_host)[1];
  const int nBig = reinterpret_cast<const int*>(g->sm_host)[3];
  if (nBig == E) {  // nothing for kernel S: the staged pipeline takes the whole call, and the next calls do not try again
    g->small_skip = 32;
    *fell_through = true;
    return TLC_OK;
  }
  cudaStream_t sb = g->small_stream ? g->small_stream : st;
  cudaStream_t sc = g->small_stream2 ? g->small_stream2 : st;
  if (nB > 0) {
    if (sb != st) CK(cudaStreamWaitEvent(sb, g->ev_small_a, 0));
    if (evb0) cudaEventRecord(evb0, sb);
    launch_small(g->gv, p, d_targets, E, vs, d_pi, d_pi32, d_status, g->sm_list_b, g->sm_list_c, g->sm_list_big, counters,
                 nullptr, nullptr, nullptr, d_stats, g->sm_count, 8, sb, nullptr, g->sm_list_b2);
    if (ev1b) cudaEventRecord(ev1b, sb);
    if (sb != st) CK(cudaEventRecord(g->ev_small_bc, sb));
  }
  if (nCc > 0) {
    if (sc != st) CK(cudaStreamWaitEvent(sc, g->ev_small_a, 0));
    if (evc0) cudaEventRecord(evc0, sc);
    launch_small(g->gv, p, d_targets, E, vs, d_pi, d_pi32, d_status, g->sm_list_b, g->sm_list_c, g->sm_list_big, counters,
                 nullptr, nullptr, nullptr, d_stats, g->sm_count, 16, sc, nullptr, g->sm_list_b2);
    if (ev2) cudaEventRecord(ev2, sc);
    if (sc != st) CK(cudaEventRecord(g->ev_small_c, sc));
  }
  StagedAcc acc;
  if (nBig > 0 && (rc = run_staged_list(g, d_targets, g->sm_list_big, nBig, up, d_pi, d_pi32, d_status, tm, acc))) return rc;
  if (nB > 0 && sb != st) CK(cudaStreamWaitEvent(st, g->ev_small_bc, 0));
  if (nCc > 0 && sc != st) CK(cudaStreamWaitEvent(st, g->ev_small_c, 0));
  CK(cudaMemcpyAsync(g->sm_host, g->sm_dev, SM_STATS_OFF + sizeof(SmallStats), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int nC = reinterpret_cast<const int*>(g->sm_host)[2];  // rows classes B / C could not take after all (listed in sm_list_b)
  const SmallStats hs = *reinterpret_cast<const SmallStats*>(g->sm_host + SM_STATS_OFF);
  float msA = 0, msB = 0, msC = 0;
  if (g->timing) {
    cudaEventElapsedTime(&msA, ev0, ev1);
    if (nB > 0) cudaEventElapsedTime(&msB, evb0, ev1b);
    if (nCc > 0) cudaEventElapsedTime(&msC, evc0, ev2);
  }
  if (nC > 0 && (rc = run_staged_list(g, d_targets, g->sm_list_b, nC, up, d_pi, d_pi32, d_status, tm, acc))) return rc;
  // a call whose vicinities are mostly too large: do not try again for a while (the size check alone costs a pass over
  // two ball bitmaps per target)
  if (((int64_t)nC + nBig) * 4 > E * 3) g->small_skip = 32;
  for (int i = 0; i < 10; i++) g->stage_ms[i] = acc.ms[i];
  g->stage_ms[9] += msA + msB + msC;
  g->small_ms[0] = msA; g->small_ms[1] = msB; g->small_ms[2] = msC;
  g->small_rows[0] = (int64_t)hs.handled[0]; g->small_rows[1] = (int64_t)hs.handled[1]; g->small_rows[2] = (int64_t)hs.handled[2];
  g->small_rows[3] = (int64_t)nC + nBig;
  g->last_live = acc.live + (int64_t)hs.live; g->last_nv = acc.nv + (int64_t)hs.sum_n; g->last_ne = acc.ne + (int64_t)hs.sum_m;
  g->last_fb = acc.fb; g->last_general = acc.general; g->last_rowcheck = acc.rowcheck; g->last_blocks = acc.blocks;
  g->last_direct = acc.direct; g->last_table = acc.table;
  g->nchunks = acc.chunks;
  g->alg_bytes = acc.bytes + hs.bytes;
  if (cnt_compute) {
    *cnt_compute = 0;
    if (d_status) {
      if ((rc = ensure_pinned(g, (size_t)E + ALIGN))) return rc;
      uint8_t* h_st = reinterpret_cast<uint8_t*>(g->h_pin);
      CK(cudaMemcpyAsync(h_st, d_status, (size_t)E, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      int64_t cnt = 0;
      for (int64_t i = 0; i < E; i++) cnt += h_st[i] <= TLC_ST_TRIVIAL;
      *cnt_compute = cnt;
    }
  }
  CK(cudaGetLastError());
  return TLC_OK;
}

// the whole path over device-resident targets: kernel S where it applies, the staged pipeline for everything else
static int run_pipeline(tlc_graph* g, const int32_t* d_targets, int64_t E, const tlc_params* up, double* d_pi,
                        float* d_pi32, uint8_t* d_status, int64_t* cnt_compute, tlc_detail* detail) {
  int rc = check_params(up);
  if (rc) return rc;
  g->small_ms[0] = g->small_ms[1] = g->small_ms[2] = 0;
  g->small_rows[0] = g->small_rows[1] = g->small_rows[2] = g->small_rows[3] = 0;
  if (small_applicable(g, up, E, detail)) {
    bool fell = false;
    rc = run_small(g, d_targets, E, up, d_pi, d_pi32, d_status, cnt_compute, &fell);
    if (rc || !fell) return rc;
  }
  return run_staged(g, d_targets, E, up, d_pi, d_pi32, d_status, cnt_compute, detail);
}

// =================================================================================================
// C-ABI
// =================================================================================================
static int graph_upload(tlc_graph* g, int32_t N, int64_t nnz, const int32_t* rowptr, const int32_t* col, const double* kappa);

extern "C" {

const char* tlc_last_error(void) { return g_err.c_str(); }
const char* tlc_version(void) { return "tlc_b200 0.1 (sm_100a)"; }
int64_t tlc_launch_count(void) { return launch_count(); }

int tlc_graph_create(int32_t N, int64_t nnz, const int32_t* rowptr, const int32_t* col, const double* kappa, int device,
                     uint64_t arena_bytes, tlc_graph** out) {
  if (!out || !rowptr || (nnz > 0 && (!col || !kappa)) || N <= 0 || nnz < 0) return fail(TLC_E_INVALID, "bad graph arguments");
  if (rowptr[0] != 0 || rowptr[N] != nnz) return fail(TLC_E_INVALID, "rowptr[0] != 0 or rowptr[N] != nnz");
  // The kernels index bitmaps / ball rows by col[] and bisect the rows: a malformed CSR would mean out-of-bounds device
  // accesses, so the whole contract is checked here, once per graph (O(nnz log deg) on the host):
  //   rowptr monotone; col in [0, N), strictly ascending inside a row (no duplicates), no self-loops; every (x, y) has its
  //   mirror (y, x) with the same curvature (graph2pi.__init__ sets ricci_curv[(a,b)] = ricci_curv[(b,a)], riccidist2dgm.py:222-225);
  //   weight kappa + 1 finite and > 0 (networkx's Dijkstra precondition; kernel 1b orders distances by their bit patterns)
  for (int32_t x = 0; x < N; x++)
    if (rowptr[x + 1] < rowptr[x]) return fail(TLC_E_INVALID, "rowptr is not monotone at node " + std::to_string(x));
  for (int32_t x = 0; x < N; x++) {  // pass 1: every row well formed on its own (pass 2 bisects the rows)
    for (int64_t e = rowptr[x]; e < rowptr[x + 1]; e++) {
      const int32_t y = col[e];
      if (y < 0 || y >= N) return fail(TLC_E_INVALID, "col out of range in row " + std::to_string(x));
      if (y == x) return fail(TLC_E_INVALID, "self-loop at node " + std::to_string(x));
      if (e > rowptr[x] && col[e - 1] >= y) return fail(TLC_E_INVALID, "row " + std::to_string(x) + " is not strictly ascending");
      const double w = kappa[e] + 1.0;
      if (!(w > 0.0) || !std::isfinite(w))
        return fail(TLC_E_INVALID, "kappa + 1 must be finite and > 0 (edge " + std::to_string(x) + "-" + std::to_string(y) + ")");
    }
  }
  for (int32_t x = 0; x < N; x++) {  // pass 2: symmetry
    for (int64_t e = rowptr[x]; e < rowptr[x + 1]; e++) {
      const int32_t y = col[e];
      const int32_t* lo = col + rowptr[y];
      const int32_t* hi = col + rowptr[y + 1];
      const int32_t* it = std::lower_bound(lo, hi, x);
      if (it == hi || *it != x) return fail(TLC_E_INVALID, "graph is not symmetric: (" + std::to_string(x) + "," + std::to_string(y) + ") has no mirror entry");
      if (kappa[it - col] != kappa[e]) return fail(TLC_E_INVALID, "curvature of (" + std::to_string(x) + "," + std::to_string(y) + ") differs from its mirror entry");
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TLC_E_NODEVICE, "no CUDA device");
  if (device < 0 || device >= ndev) return fail(TLC_E_INVALID, "device index out of range");
  CK(cudaSetDevice(device));
  tlc_graph* g = new tlc_graph();
  g->device = device;
  g->arena_req = (size_t)arena_bytes;
  const int rc = graph_upload(g, N, nnz, rowptr, col, kappa);
  if (rc != TLC_OK) { const std::string keep = g_err; tlc_graph_destroy(g); g_err = keep; return rc; }  // nothing leaks on a failed create
  *out = g;
  return TLC_OK;
}

}  // extern "C"

// device side of tlc_graph_create: every allocation is stored in *g as soon as it exists, so the caller can free a
// partially built graph with tlc_graph_destroy
static int graph_upload(tlc_graph* g, int32_t N, int64_t nnz, const int32_t* rowptr, const int32_t* col, const double* kappa) {
  const int device = g->device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  g->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&g->own_stream, cudaStreamNonBlocking));
  g->stream = g->own_stream;
  if (!getenv("TLC_NO_SIDE_STREAMS")) {
    CK(cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < 3; i++) {
      CK(cudaStreamCreateWithFlags(&g->side[i], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&g->ev_join[i], cudaEventDisableTiming));
    }
  }
  g->gv = GraphView{N, nnz, nullptr, nullptr, nullptr, nullptr};
  CK(cudaMalloc((void**)&g->gv.rowptr, (size_t)(N + 1) * 4));
  CK(cudaMalloc((void**)&g->gv.col, std::max<size_t>((size_t)nnz * 4, 16)));
  CK(cudaMalloc((void**)&g->gv.kappa, std::max<size_t>((size_t)nnz * 8, 16)));
  CK(cudaMemcpy((void*)g->gv.rowptr, rowptr, (size_t)(N + 1) * 4, cudaMemcpyHostToDevice));
  if (nnz) {
    CK(cudaMemcpy((void*)g->gv.col, col, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((void*)g->gv.kappa, kappa, (size_t)nnz * 8, cudaMemcpyHostToDevice));
  }
  {
    // interleaved row records for kernel 1b's graph-row route: {id, 0, weight}; the weight is this IEEE double add
    struct Rec { uint32_t id, pad; double w; };
    static_assert(sizeof(Rec) == 16, "row record must be 16 bytes");
    std::vector<Rec> rec((size_t)nnz);
    for (int64_t e = 0; e < nnz; e++) rec[(size_t)e] = Rec{(uint32_t)col[e], 0u, kappa[e] + 1.0};
    CK(cudaMalloc((void**)&g->gv.rec, std::max<size_t>((size_t)nnz * 16, 16)));
    if (nnz) CK(cudaMemcpy((void*)g->gv.rec, rec.data(), (size_t)nnz * 16, cudaMemcpyHostToDevice));
  }
  {
    // per node: smallest weight kappa + 1 of its row as a float rounded DOWN (settling margin of the graph-row route)
    std::vector<float> mw((size_t)N, 3.0e38f);
    for (int32_t x = 0; x < N; x++) {
      double lo = 3.0e38;
      for (int32_t e = rowptr[x]; e < rowptr[x + 1]; e++) lo = std::min(lo, kappa[e] + 1.0);
      float f = (float)lo;
      if ((double)f > lo) f = std::nextafterf(f, -INFINITY);
      mw[(size_t)x] = f;
    }
    CK(cudaMalloc((void**)&g->gminw, (size_t)N * sizeof(float)));
    CK(cudaMemcpy(g->gminw, mw.data(), (size_t)N * sizeof(float), cudaMemcpyHostToDevice));
  }

  CK(cudaMalloc((void**)&g->work_counter, 64));
  return TLC_OK;
}

extern "C" {

int tlc_graph_destroy(tlc_graph* g) {
  if (!g) return TLC_OK;
  cudaSetDevice(g->device);
  cudaFree((void*)g->gv.rowptr); cudaFree((void*)g->gv.col); cudaFree((void*)g->gv.kappa); cudaFree((void*)g->gv.rec);
  cudaFree(g->arena); cudaFree(g->bitmaps); cudaFree(g->queue); cudaFree(g->work_counter);
  cudaFree(g->ball_cache); cudaFree(g->ball_acc); cudaFree(g->ball_state); cudaFree(g->ball_list);
  cudaFree(g->d_n); cudaFree(g->d_m); cudaFree(g->d_ds); cudaFree(g->d_st); cudaFree(g->d_bytes);
  cudaFree(g->io_t); cudaFree(g->io_pi); cudaFree(g->io_st); cudaFree(g->gminw);
  if (g->h_pin) cudaFreeHost(g->h_pin);
  if (g->own_stream) cudaStreamDestroy(g->own_stream);
  for (int i = 0; i < 3; i++) { if (g->side[i]) cudaStreamDestroy(g->side[i]); if (g->ev_join[i]) cudaEventDestroy(g->ev_join[i]); }
  if (g->ev_fork) cudaEventDestroy(g->ev_fork);
  if (g->small_stream) cudaStreamDestroy(g->small_stream);
  if (g->small_stream2) cudaStreamDestroy(g->small_stream2);
  if (g->ev_small_c) cudaEventDestroy(g->ev_small_c);
  cudaFree(g->sm_list_b2);
  if (g->ev_small_a) cudaEventDestroy(g->ev_small_a);
  if (g->ev_small_bc) cudaEventDestroy(g->ev_small_bc);
  cudaFree(g->sm_list_big);
  cudaFree(g->sm_list_b); cudaFree(g->sm_list_c); cudaFree(g->sm_sub); cudaFree(g->sm_idx); cudaFree(g->sm_pi);
  cudaFree(g->sm_pi32); cudaFree(g->sm_st); cudaFree(g->sm_dev);
  if (g->sm_host) cudaFreeHost(g->sm_host);
  cudaFree(g->sssp.D); cudaFree(g->sssp.Q); cudaFree(g->sssp.P); cudaFree(g->sssp.PW); cudaFree(g->sssp.state); cudaFree(g->sssp.list);
  cudaFree(g->sssp.count); cudaFree(g->sssp_pw);
  for (void* q : g->peer_opened) cudaIpcCloseMemHandle(q);
  cudaFree(g->peer_own); cudaFree(g->peer_ticket); cudaFree(g->px_pi32); cudaFree(g->px_pi); cudaFree(g->px_st);
  for (cudaEvent_t e : g->ev_pool) cudaEventDestroy(e);
  delete g;
  return TLC_OK;
}

int tlc_vicinity_pi_dev(tlc_graph* g, const int32_t* dev_targets, int64_t E, const tlc_params* p, double* dev_out_pi,
                        float* dev_out_pi_f32, uint8_t* dev_out_status, int64_t* cnt_compute) {
  if (!g || (E > 0 && (!dev_targets || !dev_out_pi))) return fail(TLC_E_INVALID, "NULL argument");
  uint8_t* st = dev_out_status;
  if (!st && cnt_compute && E > 0) {  // the count needs the statuses: the graph's grow-only staging, no cudaMalloc per call
    CK(cudaSetDevice(g->device));
    const int rc0 = ensure_io_buffers(g, E, 0);
    if (rc0) return rc0;
    st = g->io_st;
  }
  return run_pipeline(g, dev_targets, E, p, dev_out_pi, dev_out_pi_f32, st, cnt_compute, nullptr);
}

int tlc_vicinity_pi(tlc_graph* g, const int32_t* targets, int64_t E, const tlc_params* p, double* out_pi,
                    uint8_t* out_status, int64_t* cnt_compute) {
  if (!g || (E > 0 && (!targets || !out_pi))) return fail(TLC_E_INVALID, "NULL argument");
  int rc = check_params(p);
  if (rc) return rc;
  if (cnt_compute) *cnt_compute = 0;
  if (E == 0) return TLC_OK;
  CK(cudaSetDevice(g->device));
  const int r2 = p->resolution * p->resolution;
  if ((rc = ensure_io_buffers(g, E, r2))) return rc;  // grow-only device staging: no cudaMalloc/cudaFree per call
  int32_t* d_t = g->io_t;
  double* d_pi = g->io_pi;
  uint8_t* d_st = g->io_st;
  CK(cudaMemcpyAsync(d_t, targets, (size_t)E * 8, cudaMemcpyHostToDevice, g->stream));
  // (cnt_compute is taken from the statuses that travel to the host anyway: no extra copy + synchronisation inside the pipeline)
  rc = run_pipeline(g, d_t, E, p, d_pi, nullptr, d_st, nullptr, nullptr);
  if (rc == TLC_OK) {
    uint8_t* h_st = out_status;
    if (!h_st && cnt_compute) {
      if ((rc = ensure_pinned(g, (size_t)E + ALIGN))) return rc;
      h_st = reinterpret_cast<uint8_t*>(g->h_pin);
    }
    cudaError_t e1 = cudaMemcpyAsync(out_pi, d_pi, (size_t)E * r2 * 8, cudaMemcpyDeviceToHost, g->stream);
    cudaError_t e2 = h_st ? cudaMemcpyAsync(h_st, d_st, (size_t)E, cudaMemcpyDeviceToHost, g->stream) : cudaSuccess;
    cudaError_t e3 = cudaStreamSynchronize(g->stream);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) rc = fail(TLC_E_CUDA, "copy back failed");
    else if (cnt_compute) {  // targets whose image row was really computed (riccidist2dgm.py:354)
      int64_t cnt = 0;
      for (int64_t i = 0; i < E; i++) cnt += h_st[i] <= TLC_ST_TRIVIAL;
      *cnt_compute = cnt;
    }
  }
  return rc;
}

int tlc_vicinity_sizes(tlc_graph* g, const int32_t* targets, int64_t E, const tlc_params* up, int32_t* out_n,
                       int32_t* out_m, uint8_t* out_status) {
  if (!g || (E > 0 && !targets)) return fail(TLC_E_INVALID, "NULL argument");
  int rc = check_params(up);
  if (rc) return rc;
  if (E == 0) return TLC_OK;
  CK(cudaSetDevice(g->device));
  Params p{up->hop, up->mode, up->descriptor, up->resolution, up->flags, up->img_mask};
  if ((rc = ensure_call_buffers(g, E))) return rc;
  if ((rc = ensure_vicinity_scratch(g, p))) return rc;
  if ((rc = ensure_io_buffers(g, E, 0))) return rc;  // grow-only device staging: no cudaMalloc / cudaFree per call
  int32_t* d_t = g->io_t;
  CK(cudaMemcpyAsync(d_t, targets, (size_t)E * 8, cudaMemcpyHostToDevice, g->stream));
  VicinityScratch vs = make_vs(g);
  launch_ball_cache(g->gv, p, d_t, E, vs, g->stream);
  launch_vicinity_sizes(g->gv, p, d_t, E, g->d_n, g->d_m, g->d_ds, g->d_st, g->d_bytes, vs, g->work_counter, g->stream);
  if (out_n) CK(cudaMemcpyAsync(out_n, g->d_n, (size_t)E * 4, cudaMemcpyDeviceToHost, g->stream));
  if (out_m) CK(cudaMemcpyAsync(out_m, g->d_m, (size_t)E * 4, cudaMemcpyDeviceToHost, g->stream));
  if (out_status) CK(cudaMemcpyAsync(out_status, g->d_st, (size_t)E, cudaMemcpyDeviceToHost, g->stream));
  CK(cudaStreamSynchronize(g->stream));
  CK(cudaGetLastError());
  return TLC_OK;
}

int tlc_vicinity_detail(tlc_graph* g, const int32_t* targets, int64_t E, const tlc_params* p, tlc_detail* out) {
  if (!g || !out || (E > 0 && !targets)) return fail(TLC_E_INVALID, "NULL argument");
  int rc = check_params(p);
  if (rc) return rc;
  if (E == 0) return TLC_OK;
  CK(cudaSetDevice(g->device));
  const int r2 = p->resolution * p->resolution;
  if ((rc = ensure_io_buffers(g, E, r2))) return rc;  // grow-only device staging shared with tlc_vicinity_pi
  int32_t* d_t = g->io_t;
  double* d_pi = g->io_pi;
  uint8_t* d_st = g->io_st;
  CK(cudaMemcpyAsync(d_t, targets, (size_t)E * 8, cudaMemcpyHostToDevice, g->stream));
  int64_t cnt = 0;
  g->detail_chunk.T = 0;
  rc = run_pipeline(g, d_t, E, p, d_pi, nullptr, d_st, &cnt, out);
  if (rc == TLC_OK && out->pi) {
    if (cudaMemcpy(out->pi, d_pi, (size_t)E * r2 * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(TLC_E_CUDA, "copy back failed");
  }
  // the two single-kind images of the PDGNN generators: kernel 4 again over the chunk the call left in the arena
  for (int which = 0; which < 2 && rc == TLC_OK; which++) {
    double* dst = which == 0 ? out->pi_up : out->pi_one;
    if (!dst) continue;
    Params pp{p->hop, p->mode, p->descriptor, p->resolution, p->flags, 1u << (which == 0 ? TLC_K_UP : TLC_K_ONE)};
    if (cudaMemsetAsync(d_pi, 0, (size_t)E * r2 * 8, g->stream) != cudaSuccess) { rc = fail(TLC_E_CUDA, "memset failed"); break; }
    if (g->detail_chunk.T == E) launch_pimg(pp, g->detail_chunk, d_pi, nullptr, nullptr, 64, g->stream);
    if (cudaMemcpyAsync(dst, d_pi, (size_t)E * r2 * 8, cudaMemcpyDeviceToHost, g->stream) != cudaSuccess ||
        cudaStreamSynchronize(g->stream) != cudaSuccess)
      rc = fail(TLC_E_CUDA, "copy back failed");
  }
  return rc;
}

int tlc_small_diagrams(tlc_graph* g, const int32_t* targets, int64_t E, const tlc_params* up, const int64_t* poff,
                       int32_t* npairs, int32_t* pkind, int32_t* pbv, int32_t* pdv, double* pbirth, double* pdeath,
                       double* out_pi, uint8_t* out_status, int32_t* out_n, int32_t* out_m) {
  if (!g || (E > 0 && (!targets || !poff))) return fail(TLC_E_INVALID, "NULL argument");
  int rc = check_params(up);
  if (rc) return rc;
  if (E == 0) return TLC_OK;
  if (E >= ((int64_t)1 << 31)) return fail(TLC_E_INVALID, "E must be < 2^31");
  if (up->resolution != 5 || up->descriptor < 0 || up->descriptor > 2 ||
      (up->flags & TLC_F_FILT_ANY_STRUCT))
    return fail(TLC_E_INVALID, "kernel S serves the Ricci-distance filtrations with a 5 x 5 image");
  CK(cudaSetDevice(g->device));
  Params p{up->hop, up->mode, up->descriptor, up->resolution, up->flags, up->img_mask};
  if ((rc = ensure_call_buffers(g, E))) return rc;
  if ((rc = ensure_vicinity_scratch(g, p))) return rc;
  if (!g->ball_cache) return fail(TLC_E_NOMEM, "no ball cache for this graph (TLC_BALL_CACHE_GB)");
  if ((rc = ensure_small_buffers(g, E))) return rc;
  if ((rc = ensure_io_buffers(g, E, 25))) return rc;
  const int64_t P = poff[E];
  DevBuf b_poff, b_np, b_kind, b_bv, b_dv, b_birth, b_death;
  CK(b_poff.alloc((size_t)(E + 1) * 8)); CK(b_np.alloc((size_t)E * 4)); CK(b_kind.alloc((size_t)P + 16));
  CK(b_bv.alloc((size_t)P * 4 + 16)); CK(b_dv.alloc((size_t)P * 4 + 16)); CK(b_birth.alloc((size_t)P * 8 + 16)); CK(b_death.alloc((size_t)P * 8 + 16));
  cudaStream_t st = g->stream;
  CK(cudaMemcpyAsync(g->io_t, targets, (size_t)E * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(b_poff.p, poff, (size_t)(E + 1) * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(b_np.p, 0, (size_t)E * 4, st));
  CK(cudaMemsetAsync(g->io_st, 255, (size_t)E, st));  // rows kernel S does not take keep TLC_ST_NOT_SMALL
  CK(cudaMemsetAsync(g->io_pi, 0, (size_t)E * 25 * 8, st));
  CK(cudaMemsetAsync(g->d_n, 0, (size_t)E * 4, st));
  CK(cudaMemsetAsync(g->d_m, 0, (size_t)E * 4, st));
  SmallDiag dg{b_poff.as<int64_t>(), b_np.as<int32_t>(), b_kind.as<uint8_t>(), b_bv.as<int32_t>(), b_dv.as<int32_t>(),
               b_birth.as<double>(), b_death.as<double>()};
  VicinityScratch vs = make_vs(g);
  launch_ball_cache(g->gv, p, g->io_t, E, vs, st);
  launch_small(g->gv, p, g->io_t, E, vs, g->io_pi, nullptr, g->io_st, g->sm_list_b, g->sm_list_c, g->sm_list_big,
               reinterpret_cast<int*>(g->sm_dev), g->d_n, g->d_m, &dg, nullptr, g->sm_count, 3, st, nullptr);
  std::vector<uint8_t> k8((size_t)P + 1);
  if (npairs) CK(cudaMemcpyAsync(npairs, b_np.p, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
  if (P > 0) {
    CK(cudaMemcpyAsync(k8.data(), b_kind.p, (size_t)P, cudaMemcpyDeviceToHost, st));
    if (pbv) CK(cudaMemcpyAsync(pbv, b_bv.p, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    if (pdv) CK(cudaMemcpyAsync(pdv, b_dv.p, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    if (pbirth) CK(cudaMemcpyAsync(pbirth, b_birth.p, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    if (pdeath) CK(cudaMemcpyAsync(pdeath, b_death.p, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
  }
  if (out_pi) CK(cudaMemcpyAsync(out_pi, g->io_pi, (size_t)E * 25 * 8, cudaMemcpyDeviceToHost, st));
  if (out_status) CK(cudaMemcpyAsync(out_status, g->io_st, (size_t)E, cudaMemcpyDeviceToHost, st));
  if (out_n) CK(cudaMemcpyAsync(out_n, g->d_n, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
  if (out_m) CK(cudaMemcpyAsync(out_m, g->d_m, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (pkind) for (int64_t i = 0; i < P; i++) pkind[i] = k8[(size_t)i];
  return TLC_OK;
}

int tlc_ollivier_ricci(int device, int32_t N, int64_t nnz, const int32_t* rowptr, const int32_t* col, double alpha,
                       double* out_kappa, int32_t* out_iters) {
  if (N <= 0 || nnz < 0 || !rowptr || (nnz > 0 && (!col || !out_kappa))) return fail(TLC_E_INVALID, "bad arguments");
  if (rowptr[0] != 0 || rowptr[N] != nnz) return fail(TLC_E_INVALID, "rowptr[0] != 0 or rowptr[N] != nnz");
  if (!(alpha >= 0.0 && alpha <= 1.0)) return fail(TLC_E_INVALID, "alpha must lie in [0, 1]");
  // the same structural contract as tlc_graph_create: ascending rows without duplicates or self-loops, symmetric
  for (int32_t x = 0; x < N; x++) {
    if (rowptr[x + 1] < rowptr[x]) return fail(TLC_E_INVALID, "rowptr is not monotone at node " + std::to_string(x));
    for (int64_t e = rowptr[x]; e < rowptr[x + 1]; e++) {
      const int32_t y = col[e];
      if (y < 0 || y >= N) return fail(TLC_E_INVALID, "col out of range in row " + std::to_string(x));
      if (y == x) return fail(TLC_E_INVALID, "self-loop at node " + std::to_string(x));  // (OllivierRicci removes self-loops)
      if (e > rowptr[x] && col[e - 1] >= y) return fail(TLC_E_INVALID, "row " + std::to_string(x) + " is not strictly ascending");
    }
  }
  // edge list (x < y), mirror positions, support sizes
  const int TOPK = 3000;
  std::vector<int64_t> epos, emir;
  std::vector<int32_t> esrc;
  std::vector<int64_t> esize;
  for (int32_t x = 0; x < N; x++) {
    for (int64_t e = rowptr[x]; e < rowptr[x + 1]; e++) {
      const int32_t y = col[e];
      const int32_t* lo = col + rowptr[y];
      const int32_t* hi = col + rowptr[y + 1];
      const int32_t* it = std::lower_bound(lo, hi, x);
      if (it == hi || *it != x) return fail(TLC_E_INVALID, "graph is not symmetric: (" + std::to_string(x) + "," + std::to_string(y) + ") has no mirror entry");
      if (x < y) {
        epos.push_back(e); emir.push_back(it - col); esrc.push_back(x);
        const int64_t sa = std::min<int64_t>(rowptr[x + 1] - rowptr[x], TOPK) + 1, sb = std::min<int64_t>(rowptr[y + 1] - rowptr[y], TOPK) + 1;
        esize.push_back(std::max(sa, sb));
      }
    }
  }
  if (nnz == 0) return TLC_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TLC_E_NODEVICE, "no CUDA device");
  if (device < 0 || device >= ndev) return fail(TLC_E_INVALID, "device index out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int sms = prop.multiProcessorCount;
  const int64_t E = (int64_t)epos.size();
  // size classes: both supports <= 120 (vectors and the code matrix in shared memory, several CTAs per SM),
  // <= 1024 (vectors in shared memory, codes in HBM), the rest (<= 3001)
  std::vector<int64_t> order((size_t)E);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t p, int64_t q) { return esize[(size_t)p] < esize[(size_t)q]; });
  std::vector<int64_t> h_pos((size_t)E), h_mir((size_t)E);
  std::vector<int32_t> h_src((size_t)E);
  int64_t c0 = 0, c1 = 0;
  for (int64_t i = 0; i < E; i++) {
    const int64_t o = order[(size_t)i];
    h_pos[(size_t)i] = epos[(size_t)o]; h_mir[(size_t)i] = emir[(size_t)o]; h_src[(size_t)i] = esrc[(size_t)o];
    if (esize[(size_t)o] <= 120) c0 = i + 1;
    if (esize[(size_t)o] <= 1024) c1 = i + 1;
  }
  const size_t Wc = ((size_t)N + 31) / 32;
  DevBuf b_rp, b_col, b_b1, b_b2, b_acc, b_state, b_list, b_cnt, b_tg, b_pos, b_mir, b_src, b_out, b_it, b_wsum, b_slab, b_bm;
  CK(b_rp.alloc((size_t)(N + 1) * 4)); CK(b_col.alloc((size_t)nnz * 4));
  CK(b_b1.alloc((size_t)N * Wc * 4)); CK(b_b2.alloc((size_t)N * Wc * 4));
  CK(b_acc.alloc((size_t)N * 16)); CK(b_state.alloc((size_t)N * 4)); CK(b_list.alloc((size_t)N * 4)); CK(b_cnt.alloc(64));
  CK(b_tg.alloc((size_t)N * 8));
  CK(b_pos.alloc((size_t)E * 8)); CK(b_mir.alloc((size_t)E * 8)); CK(b_src.alloc((size_t)E * 4));
  CK(b_out.alloc((size_t)nnz * 8)); CK(b_it.alloc((size_t)nnz * 4)); CK(b_wsum.alloc((size_t)(TOPK + 1) * 8));
  cudaStream_t st = nullptr;
  CK(cudaMemcpy(b_rp.p, rowptr, (size_t)(N + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b_col.p, col, (size_t)nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b_pos.p, h_pos.data(), (size_t)E * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b_mir.p, h_mir.data(), (size_t)E * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b_src.p, h_src.data(), (size_t)E * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(b_out.p, 0, (size_t)nnz * 8));
  CK(cudaMemset(b_it.p, 0, (size_t)nnz * 4));
  {
    std::vector<int32_t> tg((size_t)N * 2);
    for (int32_t i = 0; i < N; i++) { tg[2 * (size_t)i] = i; tg[2 * (size_t)i + 1] = i; }
    CK(cudaMemcpy(b_tg.p, tg.data(), (size_t)N * 8, cudaMemcpyHostToDevice));
    // neighbour weight base ** (-(w ** exp_power)) = e ** -1 and its running sums, as python's sum() adds them
    const double wnb = std::exp(-1.0);
    std::vector<double> ws((size_t)TOPK + 1, 0.0);
    for (int k = 1; k <= TOPK; k++) ws[(size_t)k] = ws[(size_t)k - 1] + wnb;
    CK(cudaMemcpy(b_wsum.p, ws.data(), ws.size() * 8, cudaMemcpyHostToDevice));
  }
  // closed 1-hop and 2-hop balls of every node (kernel 1's ball expansion)
  GraphView gv{N, nnz, b_rp.as<int32_t>(), b_col.as<int32_t>(), nullptr, nullptr};
  for (int hop = 1; hop <= 2; hop++) {
    Params pp{hop, TLC_MODE_NODE, TLC_DESC_SUM, 5, 0, 0};
    size_t Wd = 0;
    bool smem = false;
    const int grid = vicinity_grid(device, gv, pp, &Wd, &smem);
    if (!smem && !b_bm.p) CK(b_bm.alloc((size_t)grid * 2 * Wd * 4));
    CK(cudaMemset(b_state.p, 0, (size_t)N * 4));
    VicinityScratch vs{smem ? nullptr : b_bm.as<uint32_t>(), nullptr, grid, (hop == 1 ? b_b1 : b_b2).as<uint32_t>(),
                       b_acc.as<unsigned long long>(), b_state.as<int32_t>(), b_list.as<int32_t>(), b_cnt.as<int>()};
    launch_ball_cache(gv, pp, b_tg.as<int32_t>(), N, vs, st);
  }
  const double kval[4] = {std::exp(0.0 / -1e-1), std::exp(1.0 / -1e-1), std::exp(2.0 / -1e-1), std::exp(3.0 / -1e-1)};
  const double wnb = std::exp(-1.0);
  auto run = [&](int64_t lo, int64_t hi, int cap, bool codes_smem, int per_sm) -> int {
    if (hi <= lo) return TLC_OK;
    const int grid = (int)std::min<int64_t>(hi - lo, (int64_t)sms * per_sm);
    size_t stride = 0;
    if (!codes_smem) {
      stride = ((size_t)cap * cap + 255) / 256 * 256;
      if (b_slab.p) { cudaFree(b_slab.p); b_slab.p = nullptr; }
      CK(b_slab.alloc(stride * (size_t)grid));
    }
    launch_ricci(b_rp.as<int32_t>(), b_col.as<int32_t>(), b_b1.as<uint32_t>(), b_b2.as<uint32_t>(), (int)Wc,
                 b_pos.as<int64_t>() + lo, b_mir.as<int64_t>() + lo, b_src.as<int32_t>() + lo, hi - lo, alpha, kval,
                 b_wsum.as<double>(), wnb, TOPK, 1000, 1e-9, cap, codes_smem, b_slab.as<uint8_t>(), stride, grid,
                 b_out.as<double>(), b_it.as<int32_t>(), st);
    return TLC_OK;
  };
  int rc;
  if ((rc = run(0, c0, 120, true, 8))) return rc;
  if ((rc = run(c0, c1, 1024, false, 2))) return rc;
  if ((rc = run(c1, E, TOPK + 1, false, 1))) return rc;
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  CK(cudaMemcpy(out_kappa, b_out.p, (size_t)nnz * 8, cudaMemcpyDeviceToHost));
  if (out_iters) CK(cudaMemcpy(out_iters, b_it.p, (size_t)nnz * 4, cudaMemcpyDeviceToHost));
  return TLC_OK;
}

int tlc_union_find(int device, int32_t n, int32_t m, const double* fval, const int32_t* a, const int32_t* b, uint32_t flags,
                   int32_t* npairs, int32_t* pkind, int32_t* pbv, int32_t* pdv, double* pbirth, double* pdeath,
                   int32_t* npos, int32_t* pos, int32_t* nneg, int32_t* neg, uint8_t* status) {
  if (n <= 0 || m < 0 || !fval || (m > 0 && (!a || !b))) return fail(TLC_E_INVALID, "bad arguments");
  for (int32_t i = 0; i < m; i++)
    if (a[i] < 0 || a[i] >= n || b[i] < 0 || b[i] >= n) return fail(TLC_E_INVALID, "edge endpoint out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TLC_E_NODEVICE, "no CUDA device");
  CK(cudaSetDevice(device));
  char* arena = nullptr;
  const size_t bytes = chunk_bytes(1, n, m, 0);
  CK(cudaMalloc((void**)&arena, bytes));
  ChunkView c = carve(arena, 1, n, m, 0, nullptr);
  cudaStream_t st = nullptr;  // legacy default stream: the plain cudaMemcpy calls below order with it
  const int64_t zero = 0, nv = n, ne = m;
  const int32_t one_n = n, one_m = m, z32 = 0;
  const uint8_t st_ok = TLC_ST_OK;
  CK(cudaMemcpy((void*)c.tidx, &zero, 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy((void*)c.voff, &zero, 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy((void*)(c.voff + 1), &nv, 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy((void*)c.eoff, &zero, 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy((void*)(c.eoff + 1), &ne, 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.tn, &one_n, 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.tm, &one_m, 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.tnp, &z32, 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.tnpos, &z32, 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.tnneg, &z32, 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.tstatus, &st_ok, 1, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c.fval, fval, (size_t)n * 8, cudaMemcpyHostToDevice));
  if (m) {
    CK(cudaMemcpy(c.elo, a, (size_t)m * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.ehi, b, (size_t)m * 4, cudaMemcpyHostToDevice));
  }
  Params p{0, TLC_MODE_NODE, TLC_DESC_SUM, 5, flags, 0};
  const int block = block_for(m);
  launch_sort(p, c, block, 3, 0, st);
  const int64_t uf_ints = 2 * (int64_t)n * 4 <= 96 * 1024 ? 2 * (int64_t)n : 0;
  launch_union_find(p, c, block, (int)uf_ints, 1, 3, 0, st);
  if (flags & TLC_F_EXTENDED) {
    const int64_t lp_ints = 3 * (int64_t)n * 4 <= 200 * 1024 ? 3 * (int64_t)n : 0;
    launch_loops(p, c, std::min(block, 128), (int)lp_ints, n, st);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  int32_t h_np = 0, h_npos = 0, h_nneg = 0;
  uint8_t h_st = 0;
  CK(cudaMemcpy(&h_np, c.tnp, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&h_npos, c.tnpos, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&h_nneg, c.tnneg, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&h_st, c.tstatus, 1, cudaMemcpyDeviceToHost));
  if (npairs) *npairs = h_np;
  if (npos) *npos = h_npos;
  if (nneg) *nneg = h_nneg;
  if (status) *status = h_st;
  if (h_np > 0) {
    if (pkind) {
      std::vector<uint8_t> k8((size_t)h_np);
      CK(cudaMemcpy(k8.data(), c.pkind, (size_t)h_np, cudaMemcpyDeviceToHost));
      for (int i = 0; i < h_np; i++) pkind[i] = k8[i];
    }
    if (pbv) CK(cudaMemcpy(pbv, c.pbv, (size_t)h_np * 4, cudaMemcpyDeviceToHost));
    if (pdv) CK(cudaMemcpy(pdv, c.pdv, (size_t)h_np * 4, cudaMemcpyDeviceToHost));
    if (pbirth) CK(cudaMemcpy(pbirth, c.pbirth, (size_t)h_np * 8, cudaMemcpyDeviceToHost));
    if (pdeath) CK(cudaMemcpy(pdeath, c.pdeath, (size_t)h_np * 8, cudaMemcpyDeviceToHost));
  }
  if (pos && h_npos > 0) CK(cudaMemcpy(pos, c.pos, (size_t)h_npos * 4, cudaMemcpyDeviceToHost));
  if (neg && h_nneg > 0) CK(cudaMemcpy(neg, c.neg, (size_t)h_nneg * 4, cudaMemcpyDeviceToHost));
  cudaFree(arena);
  return TLC_OK;
}

int tlc_pimg_transform(int device, const double* dgm, int64_t K, int32_t resolution, double* out) {
  if (!out || (K > 0 && !dgm) || resolution < 1 || resolution > 16) return fail(TLC_E_INVALID, "bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TLC_E_NODEVICE, "no CUDA device");
  CK(cudaSetDevice(device));
  double *d_dgm = nullptr, *d_out = nullptr;
  CK(cudaMalloc((void**)&d_dgm, std::max<size_t>((size_t)K * 16, 16)));
  CK(cudaMalloc((void**)&d_out, (size_t)resolution * resolution * 8));
  if (K) CK(cudaMemcpy(d_dgm, dgm, (size_t)K * 16, cudaMemcpyHostToDevice));
  launch_pimg_single(d_dgm, K, resolution, d_out, nullptr);
  CK(cudaMemcpy(out, d_out, (size_t)resolution * resolution * 8, cudaMemcpyDeviceToHost));
  CK(cudaGetLastError());
  cudaFree(d_dgm); cudaFree(d_out);
  return TLC_OK;
}

int tlc_pi_gather(int device, const double* dev_table, int64_t rows, int32_t r2, const int64_t* dev_index, int64_t start,
                  int64_t n, float* dev_out_f32, void* stream) {
  if (rows < 0 || r2 < 1 || n < 0 || (n > 0 && (!dev_table || !dev_out_f32))) return fail(TLC_E_INVALID, "bad arguments");
  if (n == 0) return TLC_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TLC_E_NODEVICE, "no CUDA device");
  if (device < 0 || device >= ndev) return fail(TLC_E_INVALID, "device index out of range");
  CK(cudaSetDevice(device));
  // per host thread and device: the out-of-range flag and the SM count; released when the thread exits
  struct GatherScratch {
    int* d_bad[64] = {nullptr};
    int sms[64] = {0};
    ~GatherScratch() { for (int i = 0; i < 64; i++) if (d_bad[i]) cudaFree(d_bad[i]); }
  };
  static thread_local GatherScratch gs;
  int** d_bad = gs.d_bad;
  int* sms = gs.sms;
  if (device >= 64) return fail(TLC_E_INVALID, "device index out of range");
  if (!d_bad[device]) {
    CK(cudaMalloc((void**)&d_bad[device], sizeof(int)));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    sms[device] = prop.multiProcessorCount;
  }
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(d_bad[device], 0, sizeof(int), st));
  launch_gather_rows(dev_table, rows, r2, dev_index, start, n, dev_out_f32, d_bad[device], sms[device], st);
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad[device], sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (bad) return fail(TLC_E_INVALID, "row index out of range");
  return TLC_OK;
}

// ---- multi-GPU exchange by peer stores ----
int tlc_table_create(tlc_graph* g, int64_t rows, int32_t r2, void** dev_rows, unsigned char* handle64) {
  if (!g || rows <= 0 || r2 < 1 || !handle64) return fail(TLC_E_INVALID, "bad arguments");
  CK(cudaSetDevice(g->device));
  if (g->peer_own) return fail(TLC_E_INVALID, "this graph already owns an exchange table");
  const size_t bytes = PEER_HEADER_BYTES + (size_t)rows * (size_t)(r2 + 1) * sizeof(float);
  CK(cudaMalloc(&g->peer_own, bytes));
  CK(cudaMemset(g->peer_own, 0, bytes));
  if (!g->peer_ticket) { CK(cudaMalloc((void**)&g->peer_ticket, 64)); CK(cudaMemset(g->peer_ticket, 0, 64)); }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, g->peer_own));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  memcpy(handle64, &h, 64);
  g->peer_rows = rows; g->peer_r2 = r2; g->peer_epoch = 0;
  g->peers.n = 0;
  if (dev_rows) *dev_rows = reinterpret_cast<unsigned char*>(g->peer_own) + PEER_HEADER_BYTES;
  return TLC_OK;
}

int tlc_table_attach(tlc_graph* g, int32_t nranks, int32_t my_rank, const unsigned char* handles64) {
  if (!g || !g->peer_own || nranks < 1 || nranks > PEER_MAX || my_rank < 0 || my_rank >= nranks || !handles64)
    return fail(TLC_E_INVALID, "bad arguments (create the table first; at most 16 ranks)");
  CK(cudaSetDevice(g->device));
  for (void* q : g->peer_opened) cudaIpcCloseMemHandle(q);
  g->peer_opened.clear();
  g->peers.n = nranks;
  for (int r = 0; r < nranks; r++) {
    if (r == my_rank) { g->peers.table[r] = g->peer_own; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles64 + (size_t)r * 64, 64);
    void* q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    g->peer_opened.push_back(q);
    g->peers.table[r] = q;
  }
  return TLC_OK;
}

int tlc_vicinity_pi_exchange(tlc_graph* g, const int32_t* dev_targets, const int64_t* dev_row_index, int64_t E,
                             const tlc_params* p, int64_t* cnt_compute) {
  if (!g || (E > 0 && (!dev_targets || !dev_row_index))) return fail(TLC_E_INVALID, "NULL argument");
  if (!g->peer_own || g->peers.n < 1) return fail(TLC_E_INVALID, "no exchange table attached (tlc_table_create / tlc_table_attach)");
  int rc = check_params(p);
  if (rc) return rc;
  const int r2 = p->resolution * p->resolution;
  if (r2 != g->peer_r2) return fail(TLC_E_INVALID, "resolution differs from the exchange table's");
  CK(cudaSetDevice(g->device));
  if (E > g->px_cap) {
    cudaFree(g->px_pi32); cudaFree(g->px_pi); cudaFree(g->px_st);
    g->px_pi32 = nullptr; g->px_pi = nullptr; g->px_st = nullptr; g->px_cap = 0;
    const int64_t cap = E + E / 8 + 256;
    CK(cudaMalloc((void**)&g->px_pi32, (size_t)cap * r2 * 4));
    CK(cudaMalloc((void**)&g->px_pi, (size_t)cap * r2 * 8));
    CK(cudaMalloc((void**)&g->px_st, (size_t)cap));
    g->px_cap = cap;
  }
  if (E > 0) {
    rc = run_pipeline(g, dev_targets, E, p, g->px_pi, g->px_pi32, g->px_st, cnt_compute, nullptr);
    if (rc) return rc;
  } else if (cnt_compute) *cnt_compute = 0;
  // store this rank's rows into every rank's table at their final index, then wait for everyone's arrival
  launch_peer_scatter(g->px_pi32, g->px_st, dev_row_index, E, r2, g->peers, g->peer_ticket, g->sm_count, g->stream);
  g->peer_epoch++;
  launch_peer_wait(reinterpret_cast<const unsigned int*>(g->peer_own), g->peer_epoch * (unsigned int)g->peers.n, g->stream);
  CK(cudaGetLastError());
  return TLC_OK;
}

int tlc_graph_set_hks_time(tlc_graph* g, double t) {
  if (!g) return fail(TLC_E_INVALID, "NULL graph");
  if (!(t > 0.0 && t <= 50.0)) return fail(TLC_E_INVALID, "hks time must lie in (0, 50]");
  g->hks_time = t;
  return TLC_OK;
}

int tlc_graph_set_stream(tlc_graph* g, void* stream) {
  if (!g) return fail(TLC_E_INVALID, "NULL graph");
  g->stream = stream ? (cudaStream_t)stream : g->own_stream;
  return TLC_OK;
}

int tlc_last_counts(tlc_graph* g, int64_t* out5) {  // out5: 8 slots
  if (!g || !out5) return TLC_E_INVALID;
  out5[0] = g->last_live; out5[1] = g->last_nv; out5[2] = g->last_ne; out5[3] = g->nchunks; out5[4] = g->last_fb;
  out5[5] = g->last_general; out5[6] = g->last_rowcheck; out5[7] = g->last_blocks;
  return TLC_OK;
}

int64_t tlc_last_direct(tlc_graph* g) { return g ? g->last_direct : 0; }

int64_t tlc_last_table(tlc_graph* g) { return g ? g->last_table : 0; }
double tlc_table_build_ms(tlc_graph* g) { return g ? g->sssp_build_ms : 0.0; }

int tlc_last_small(tlc_graph* g, double* out7) {
  if (!g || !out7) return TLC_E_INVALID;
  for (int i = 0; i < 3; i++) out7[i] = g->small_ms[i];
  for (int i = 0; i < 4; i++) out7[3 + i] = (double)g->small_rows[i];
  return TLC_OK;
}

int tlc_last_stage_ms(tlc_graph* g, double* out10) {
  if (!g || !out10) return 0;
  for (int i = 0; i < 10; i++) out10[i] = g->stage_ms[i];
  return g->nchunks;
}

int tlc_last_algorithmic_bytes(tlc_graph* g, double* bytes_total, double* bytes_bfs, double* bytes_uf) {
  if (!g) return TLC_E_INVALID;
  if (bytes_total) *bytes_total = g->alg_bytes;
  if (bytes_bfs) *bytes_bfs = g->alg_bytes_bfs;
  if (bytes_uf) *bytes_uf = g->alg_bytes_uf;
  return TLC_OK;
}

}  // extern "C"
