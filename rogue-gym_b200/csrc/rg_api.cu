// rg_api.cu — the extern "C" boundary declared in include/rogue_b200.h.
//
// Host side of the batch driver: what ThreadConductor (python/src/thread_impls.rs:14-88) and
// GameStateImpl (python/src/state_impls.rs) do with N threads and 2N channels is here one
// device arena, one stream and one kernel launch per call. There is no CPU implementation
// behind these entry points: without a CUDA device every compute call returns RG_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/rogue_b200.h"
#include "rg_launch.h"
#include "rg_types.h"

using rg::DevBatch;
using rg::EnvState;

struct rg_batch {
  rg_params P;
  int64_t n = 0;
  int64_t max_steps = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t bg[2] = {nullptr, nullptr};  // background streams: k_prefetch passes alternate, so two can be in flight
  cudaStream_t side = nullptr;    // full-path steps, forked after the player kernel and joined at the end of the step
  cudaStream_t mir = nullptr;     // first host-mirror pass of a step
  cudaStream_t mon = nullptr;     // player + monster kernels of the envs with an active monster, beside k_step_fast
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_mir = nullptr, ev_mon = nullptr, ev_fast = nullptr;
  rg::StepStreams step_streams() const {
    rg::StepStreams q;
    q.main = stream; q.side = side; q.mon = mon; q.mir = mir;
    q.ev_fork = ev_fork; q.ev_join = ev_join; q.ev_mon = ev_mon; q.ev_fast = ev_fast; q.ev_mir = ev_mir;
    return q;
  }
  // background builder of next-level skeletons (speculative descents): one pass per step, on SPEC_STREAMS rotating
  // streams so that a pass (one skeleton ~ 140 us) has several steps' time before the main stream waits for it
  static constexpr int SPEC_STREAMS = 4;
  cudaStream_t spec[SPEC_STREAMS] = {};
  cudaEvent_t ev_spec[SPEC_STREAMS] = {};  // end of the last pass on each stream
  cudaEvent_t ev_step = nullptr;  // "this step is queued up to here" for the skeleton pass
  int64_t spec_passes = 0;
  cudaEvent_t ev_main = nullptr;  // "this step's kernels are queued up to here"
  cudaEvent_t ev_bg[2] = {nullptr, nullptr};  // end of the last pass on each background stream
  int64_t passes = 0;             // background passes kicked so far
  bool prefetch_running = false;
  int prefetch_every = 2;   // kick the background pass every k-th auto-reset step (measured: 1 -> 0.259 ms, 2 -> 0.251 ms, 4 -> 0.272 ms per step)
  int prefetch_warps = 0;   // grid-stride warps of k_prefetch (0 = as many as the full-path kernel)
  int64_t auto_steps = 0;
  int64_t steps_launched = 0;
  bool use_graph = true;
  cudaGraphExec_t graph[2] = {nullptr, nullptr};    // [auto_reset]
  cudaGraphExec_t graph_m[2] = {nullptr, nullptr};  // the same step with the two host-mirror passes
  rg::MirrorArgs margs{};
  DevBatch d{};
  rg_params* dP = nullptr;
  uint8_t* d_actions = nullptr;
  uint8_t* d_hist_bytes = nullptr;  // lazily allocated [N][C]
  uint64_t* d_u64 = nullptr;        // [2N] scratch: seeds / hashes
  int* d_out3 = nullptr;
  uint32_t* h_errflag = nullptr;    // pinned
  uint8_t* h_actions = nullptr;     // pinned + mapped [N]: staging for rg_step_mirror (k_step_scan reads it over PCIe)
  uint8_t* h_actions_dev = nullptr; // the same buffer as the device addresses it
  uint32_t* m_ticket = nullptr;     // device: see MirrorArgs::ticket
  int32_t* d_level_seen = nullptr;  // [N] deepest level rg_step_train has seen per env (lazily allocated)
  uint8_t* h_error = nullptr;       // pinned [N]
  // host mirror (rg_mirror_get): pinned + mapped host block, its device alias, and the shadows
  void* m_host = nullptr;
  size_t m_bytes = 0;
  rg_host_obs m_obs{};          // host pointers into m_host
  rg_host_obs m_dev{};          // the same addresses as the device sees them
  uint8_t* m_hist_host = nullptr;
  uint8_t* m_hist_dev = nullptr;
  uint8_t* ms_screen = nullptr;
  uint8_t* ms_hist = nullptr;
  uint8_t* ms_flat = nullptr;
  unsigned long long* m_count = nullptr;  // device counter of bytes stored to the host
  uint64_t* h_count = nullptr;            // pinned [2]: bytes stored by the last call; publishing passes completed
  unsigned long long* m_seq = nullptr;    // device: publishing passes completed
  uint64_t m_seq_expected = 0;            // publishing passes launched
  bool mirror_spin = true;                // wait for a pass by spinning on h_count[1] (RG_MIRROR_SPIN=0: cudaStreamSynchronize)
  int sm_count = 148;
  std::vector<void*> dev_allocs;
  std::string err;
  int64_t launches = 0;
  int step_parity = 0;
};

namespace {

thread_local std::string g_last_error;

int set_err(rg_batch* b, int code, const std::string& m) {
  if (b) b->err = m;
  g_last_error = m;
  return code;
}
int cuda_fail(rg_batch* b, cudaError_t e, const char* what) {
  return set_err(b, RG_ERR_CUDA, std::string("CUDA failure in ") + what + ": " + cudaGetErrorString(e));
}
#define RG_CUDA(b, call)                                  \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(b, e__, #call); \
  } while (0)

template <class T>
cudaError_t dev_alloc(rg_batch* b, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(count * sizeof(T), 16));
  if (e != cudaSuccess) return e;
  b->dev_allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return cudaSuccess;
}

const char* status_text(int code) {
  switch (code) {
    case RG_ERR_INVALID_INPUT: return "Invliad input key";
    case RG_ERR_IGNORED_INPUT: return "Ignored input code";
    case RG_ERR_PANIC: return "worker panicked (a state in which the reference implementation panics)";
    case RG_ERR_SETTING: return "Invalid Setting / InvalidTileError";
    default: return "unknown error";
  }
}

bool same_except_seed(rg_params a, rg_params b) {
  a.has_seed = b.has_seed = 0;
  a.seed_lo = b.seed_lo = a.seed_hi = b.seed_hi = 0;
  return memcmp(&a, &b, sizeof(a)) == 0;
}
// What the HBM layout and the observation shape depend on: these must agree across one batch.
bool same_geometry(const rg_params& a, const rg_params& b) {
  return a.width == b.width && a.height == b.height && a.room_num_x == b.room_num_x && a.room_num_y == b.room_num_y &&
         a.symbols == b.symbols;
}

int invalidate_prefetched(rg_batch* b);

int create_impl(const rg_params& P, const std::vector<rg_params>* per_env, int64_t n_envs, int64_t max_steps, int device,
                rg_batch** out) {
  if (!out) return set_err(nullptr, RG_ERR_ARG, "rg_create: out is null");
  *out = nullptr;
  if (n_envs < 1) return set_err(nullptr, RG_ERR_ARG, "rg_create: n_envs must be >= 1");
  if (max_steps < 0) return set_err(nullptr, RG_ERR_ARG, "rg_create: max_steps must be >= 0");
  char ebuf[256];
  int rc = rg_validate_params(&P, ebuf, sizeof(ebuf));
  if (rc != RG_OK) return set_err(nullptr, rc, ebuf);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return set_err(nullptr, RG_ERR_CUDA,
                   std::string("no CUDA device available (this library has no CPU path): ") +
                       (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0"));
  if (device < 0 || device >= ndev) return set_err(nullptr, RG_ERR_ARG, "rg_create: invalid device ordinal");
  rg_batch* b = new rg_batch();
  b->P = P;
  b->n = n_envs;
  b->max_steps = max_steps;
  b->device = device;
  auto fail = [&](int code) {
    std::string keep = b->err;
    rg_destroy(b);
    g_last_error = keep;
    return code;
  };
#define RG_TRY(call)                                                 \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return fail(cuda_fail(b, e__, #call));   \
  } while (0)
  RG_TRY(cudaSetDevice(device));
  {  // RG_MAIN_PRIO=h: the batch's stream (scan, thread-per-env kernel, its leftovers) at the branches' priority
    int lo = 0, hi = 0;
    RG_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const char* mp = getenv("RG_MAIN_PRIO");
    RG_TRY(cudaStreamCreateWithPriority(&b->stream, cudaStreamNonBlocking, (mp && mp[0] == 'h') ? hi : lo));
  }
  DevBatch& d = b->d;
  d.n = n_envs;
  d.W = P.width;
  d.H = P.height;
  d.C = d.W * d.H;
  d.CP = (d.C + 15) / 16 * 16;
  d.HB = ((d.CP / 8) + 15) / 16 * 16;
  d.WW = (d.W + 31) / 32;
  d.max_steps = max_steps;
  d.nx = P.room_num_x;
  d.ny = P.room_num_y;
  d.rsx = d.W / d.nx;
  d.rsy = d.H / d.ny;
  const size_t N = (size_t)n_envs;
  {
    cudaDeviceProp prop;
    RG_TRY(cudaGetDeviceProperties(&prop, device));
    b->sm_count = prop.multiProcessorCount;
    d.gen_warps = prop.multiProcessorCount * 16;  // grid-stride warps of the full-path kernel
    d.mon_warps = prop.multiProcessorCount * 16;  // grid-stride warps of the monster kernel
    if (const char* e = getenv("RG_PLAYER_BLOCKS")) d.player_blocks = atoi(e);
  }
  {  // per-env configs (python/src/lib.rs:270-280 builds every worker from its own JSON): the distinct
     // parameter sets go into a table, every env carries an index into it
    std::vector<rg_params> table(1, P);
    std::vector<uint16_t> idx;
    if (per_env) {
      idx.assign(N, 0);
      for (size_t i = 0; i < N; ++i) {
        const rg_params& q = (*per_env)[i];
        size_t k = 0;
        while (k < table.size() && !same_except_seed(table[k], q)) ++k;
        if (k == table.size()) {
          if (table.size() >= 65535) return fail(set_err(b, RG_ERR_SETTING, "too many distinct configs in one batch"));
          table.push_back(q);
        }
        idx[i] = (uint16_t)k;
      }
    }
    RG_TRY(dev_alloc(b, &b->dP, table.size()));
    RG_TRY(cudaMemcpy(b->dP, table.data(), table.size() * sizeof(rg_params), cudaMemcpyHostToDevice));
    d.P = b->dP;
    if (table.size() > 1) {
      uint16_t* di = nullptr;
      RG_TRY(dev_alloc(b, &di, N));
      RG_TRY(cudaMemcpy(di, idx.data(), N * sizeof(uint16_t), cudaMemcpyHostToDevice));
      d.cfg_idx = di;
    }
  }
  {  // Floor::cd_to_room_id (floor.rs:194-200) as a table: the sector grid of rooms.rs:176,191-206
    uint8_t lut[208];
    memset(lut, 0xFF, sizeof(lut));
    for (int x = 0; x < d.W; ++x)
      if (x / d.rsx < d.nx) lut[x] = (uint8_t)(x / d.rsx);
    for (int yi = 0; yi < d.ny; ++yi) {
      int ry = d.rsy, ay0;
      if (yi == 0) {
        ry -= 1;
        ay0 = 1;
      } else {
        ay0 = ry * yi;
      }
      if (ay0 + ry == d.H) ry -= 1;
      for (int y = ay0; y < ay0 + ry && y < d.H; ++y) lut[160 + y] = (uint8_t)yi;
    }
    uint8_t* dl = nullptr;
    RG_TRY(dev_alloc(b, &dl, sizeof(lut)));
    RG_TRY(cudaMemcpy(dl, lut, sizeof(lut), cudaMemcpyHostToDevice));
    d.room_lut = dl;
  }
  RG_TRY(dev_alloc(b, &d.surface, N * d.CP));
  RG_TRY(dev_alloc(b, &d.attr, N * d.CP));
  RG_TRY(dev_alloc(b, &d.screen, N * d.CP));
  RG_TRY(dev_alloc(b, &d.hist, N * d.HB));
  RG_TRY(dev_alloc(b, &d.walk, N * (size_t)d.H * d.WW));
  RG_TRY(dev_alloc(b, &d.dist, N * (size_t)rg::NCACHE * d.CP));
  RG_TRY(dev_alloc(b, &d.bfs, N * (size_t)rg::NCACHE * 2 * d.H * d.WW));
  RG_TRY(dev_alloc(b, &d.wsnap, N * (size_t)rg::NCACHE * d.H * d.WW));
  RG_TRY(dev_alloc(b, &d.st, N));
  // (+ one 64-byte line each: the host mirror reads these arrays in whole lines)
  RG_TRY(dev_alloc(b, &d.status, N * 10 + 16));
  RG_TRY(dev_alloc(b, &d.reward, N + 16));
  RG_TRY(dev_alloc(b, &d.done, N + 64));
  RG_TRY(dev_alloc(b, &d.message, N + 16));
  RG_TRY(dev_alloc(b, &d.error, N + 64));
  RG_TRY(dev_alloc(b, &d.errflag, 1));
  RG_TRY(dev_alloc(b, &d.scr_rows, N));
  RG_TRY(cudaMemsetAsync(d.scr_rows, 0xFF, N * 8, b->stream));
  RG_TRY(dev_alloc(b, &d.defer_list, N));
  RG_TRY(dev_alloc(b, &d.defer_count, 4));
  RG_TRY(dev_alloc(b, &d.full_path, N));
  RG_TRY(cudaMemsetAsync(d.full_path, 0, N, b->stream));
  RG_TRY(cudaMemsetAsync(d.defer_count, 0, 16, b->stream));
  {  // next-episode buffers + background stream (RG_PREFETCH=0 turns the pipeline off)
    const char* pf = getenv("RG_PREFETCH");
    d.prefetch = (pf && pf[0] == '0') ? 0 : 1;
    if (const char* e = getenv("RG_PREFETCH_EVERY")) b->prefetch_every = std::max(1, atoi(e));
    if (const char* e = getenv("RG_PREFETCH_WARPS")) b->prefetch_warps = std::max(32, atoi(e));
    if (d.prefetch) {
      const size_t NS = N * rg::SP_DEPTH;
      RG_TRY(dev_alloc(b, &d.sp_surface, NS * d.CP));
      RG_TRY(dev_alloc(b, &d.sp_attr, NS * d.CP));
      RG_TRY(dev_alloc(b, &d.sp_screen, NS * d.CP));
      RG_TRY(dev_alloc(b, &d.sp_hist, NS * d.HB));
      RG_TRY(dev_alloc(b, &d.sp_walk, NS * (size_t)d.H * d.WW));
      RG_TRY(dev_alloc(b, &d.sp_st, NS));
      RG_TRY(dev_alloc(b, &d.sp_state, NS));
      RG_TRY(cudaMemsetAsync(d.sp_state, 0, NS, b->stream));
      uint32_t cap = 1;
      while (cap < 4 * N) cap <<= 1;
      d.refill_cap = cap;
      RG_TRY(dev_alloc(b, &d.refill_ring, cap));
      RG_TRY(dev_alloc(b, &d.refill_ctl, 4));
      int lo = 0, hi = 0;
      RG_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      RG_TRY(dev_alloc(b, &d.refill_win, 16));
      RG_TRY(cudaMemsetAsync(d.refill_win, 0, 64, b->stream));
      RG_TRY(dev_alloc(b, &d.sp_lock, N));
      RG_TRY(dev_alloc(b, &d.sp_cancel, NS));
      d.prefetch_every = b->prefetch_every;
      if (const char* e = getenv("RG_PF_WPB")) d.pf_wpb = std::max(1, atoi(e));
      if (const char* e = getenv("RG_PF_EXCLUSIVE")) d.pf_exclusive = atoi(e) != 0;
      {
        // a pass is a few hundred one-warp chains: at high priority it starts at once and costs the
        // step kernels next to nothing; at low priority it only runs in their gaps and falls behind
        const char* pr = getenv("RG_BG_PRIO");
        for (int i = 0; i < 2; ++i) {
          RG_TRY(cudaStreamCreateWithPriority(&b->bg[i], cudaStreamNonBlocking, (pr && pr[0] == 'l') ? lo : hi));
          RG_TRY(cudaEventCreateWithFlags(&b->ev_bg[i], cudaEventDisableTiming));
        }
      }
      RG_TRY(cudaEventCreateWithFlags(&b->ev_main, cudaEventDisableTiming));
    }
  }
  if (const char* tr = getenv("RG_TRACE"); tr && tr[0] == '1') {
    RG_TRY(dev_alloc(b, &d.trace, 512 * rg::TRACE_KERNELS * 2));
    std::vector<unsigned long long> init(512 * rg::TRACE_KERNELS * 2);
    for (size_t i = 0; i < init.size(); ++i) init[i] = (i & 1) ? 0ull : ~0ull;
    RG_TRY(cudaMemcpy(d.trace, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
  }
  {  // speculative descents (RG_SPEC=0 turns them off)
    const char* sp = getenv("RG_SPEC");
    d.spec = (sp && sp[0] == '0') ? 0 : 1;
    if (d.spec) {
      RG_TRY(dev_alloc(b, &d.spec_S, N * d.CP));
      RG_TRY(dev_alloc(b, &d.spec_A, N * d.CP));
      RG_TRY(dev_alloc(b, &d.spec_rooms, N * rg::MAX_ROOMS));
      RG_TRY(dev_alloc(b, &d.spec_tag, N));
      RG_TRY(cudaMemsetAsync(d.spec_tag, 0, N * sizeof(rg::SpecTag), b->stream));
      RG_TRY(dev_alloc(b, &d.spec_seq, N));
      RG_TRY(cudaMemsetAsync(d.spec_seq, 0, N * 4, b->stream));
      RG_TRY(dev_alloc(b, &d.spec_lock, N));
      RG_TRY(cudaMemsetAsync(d.spec_lock, 0, N * 4, b->stream));
      uint32_t cap = 1;
      while (cap < 2 * N) cap <<= 1;
      d.spec_cap = cap;
      RG_TRY(dev_alloc(b, &d.spec_ring, cap));
      RG_TRY(dev_alloc(b, &d.spec_ctl, 4));
      RG_TRY(cudaMemsetAsync(d.spec_ctl, 0, 16, b->stream));
      RG_TRY(dev_alloc(b, &d.spec_win, 16));
      RG_TRY(cudaMemsetAsync(d.spec_win, 0, 64, b->stream));
      d.spec_warps = (int)std::min<int64_t>(256, ((int64_t)N + 7) / 8 * 8);
      if (const char* e = getenv("RG_SPEC_WARPS")) d.spec_warps = std::max(8, atoi(e) / 8 * 8);
      int lo = 0, hi = 0;
      RG_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      for (int i = 0; i < rg_batch::SPEC_STREAMS; ++i) {
        RG_TRY(cudaStreamCreateWithPriority(&b->spec[i], cudaStreamNonBlocking, hi));
        RG_TRY(cudaEventCreateWithFlags(&b->ev_spec[i], cudaEventDisableTiming));
      }
      RG_TRY(cudaEventCreateWithFlags(&b->ev_step, cudaEventDisableTiming));
    }
  }
  RG_TRY(dev_alloc(b, &d.stats, 8));
  RG_TRY(cudaMemsetAsync(d.stats, 0, 64, b->stream));
  RG_TRY(dev_alloc(b, &d.dstep, 4));
  RG_TRY(cudaMemsetAsync(d.dstep, 0, 16, b->stream));
  if (const char* g = getenv("RG_GRAPH")) b->use_graph = g[0] != '0';
  RG_TRY(dev_alloc(b, &d.reset_list, N));
  RG_TRY(dev_alloc(b, &d.reset_count, 4));
  RG_TRY(cudaMemsetAsync(d.reset_count, 0, 16, b->stream));
  {  // the full-path kernel's few long chains must not queue behind the player kernel's 65 536 blocks
    int lo = 0, hi = 0;
    RG_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    RG_TRY(cudaStreamCreateWithPriority(&b->side, cudaStreamNonBlocking, hi));
    RG_TRY(cudaStreamCreateWithPriority(&b->mon, cudaStreamNonBlocking, hi));
    // the first mirror pass is a light kernel that has to START early (right after k_step_fast) to hide its PCIe
    // writes behind the rest of the step: high priority, or its blocks queue behind the register-hungry warp kernels
    {
      const char* mp = getenv("RG_MIRROR_PRIO");
      RG_TRY(cudaStreamCreateWithPriority(&b->mir, cudaStreamNonBlocking, (mp && mp[0] == 'l') ? lo : hi));
      if (const char* e = getenv("RG_MIRROR_BLOCKS")) d.mirror_blocks = std::max(1, atoi(e));
    }
    RG_TRY(cudaEventCreateWithFlags(&b->ev_mir, cudaEventDisableTiming));
    RG_TRY(cudaEventCreateWithFlags(&b->ev_mon, cudaEventDisableTiming));
    RG_TRY(cudaEventCreateWithFlags(&b->ev_fast, cudaEventDisableTiming));
  }
  RG_TRY(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
  RG_TRY(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
  RG_TRY(dev_alloc(b, &d.mon_list, N));
  RG_TRY(dev_alloc(b, &d.mon_list_b, N));
  RG_TRY(dev_alloc(b, &d.mon_count, 8));
  RG_TRY(cudaMemsetAsync(d.mon_count, 0, 32, b->stream));
  RG_TRY(dev_alloc(b, &d.fast_list_m, N));
  RG_TRY(dev_alloc(b, &d.fast_list_l, N));
  RG_TRY(dev_alloc(b, &d.fast_count, 8));
  RG_TRY(cudaMemsetAsync(d.fast_count, 0, 32, b->stream));
  RG_TRY(dev_alloc(b, &d.slow_list, N));
  RG_TRY(dev_alloc(b, &d.slow_list_b, N));
  RG_TRY(dev_alloc(b, &d.slow_count, 8));
  RG_TRY(cudaMemsetAsync(d.slow_count, 0, 32, b->stream));
  d.fast = 1;
  if (const char* e = getenv("RG_FAST")) d.fast = e[0] != '0';
  if (const char* e = getenv("RG_PANIC_POLICY")) d.panic_policy = (e[0] == 't' || e[0] == '1') ? 1 : 0;
  d.branches = 1;
  if (const char* e = getenv("RG_BRANCHES")) d.branches = e[0] != '0';
  if (const char* e = getenv("RG_MON_WARPS")) d.mon_warps = std::max(32, atoi(e));
  RG_TRY(dev_alloc(b, &b->d_actions, N));
  RG_TRY(dev_alloc(b, &b->d_u64, 2 * N));
  RG_TRY(dev_alloc(b, &b->d_out3, 4));
  RG_TRY(cudaHostAlloc(&b->h_errflag, sizeof(uint32_t), cudaHostAllocMapped));
  RG_TRY(cudaHostAlloc(&b->h_actions, N, cudaHostAllocMapped));
  RG_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&b->h_actions_dev), b->h_actions, 0));
  RG_TRY(cudaMallocHost(&b->h_error, N));
  RG_TRY(cudaMemsetAsync(d.st, 0, N * sizeof(EnvState), b->stream));
  RG_TRY(cudaMemsetAsync(d.screen, ' ', N * d.CP, b->stream));
  RG_TRY(cudaMemsetAsync(d.hist, 0, N * d.HB, b->stream));
  RG_TRY(cudaMemsetAsync(d.error, 0, N, b->stream));
  RG_TRY(cudaMemsetAsync(d.errflag, 0, sizeof(uint32_t), b->stream));
  RG_TRY(rg::configure_kernels(d));
  {  // seeds: config seed (fixed, every episode identical) or fresh entropy (seed: null)
    std::vector<uint64_t> lo(N), hi(N, 0);
    bool seeded = P.has_seed != 0;
    if (per_env) {
      std::random_device rd;
      for (size_t i = 0; i < N; ++i) {
        const rg_params& q = (*per_env)[i];
        if (q.has_seed != P.has_seed)
          return fail(set_err(b, RG_ERR_SETTING, "per-env configs must all set or all omit `seed`"));
        lo[i] = q.has_seed ? q.seed_lo : (((uint64_t)rd() << 32) | rd());
        hi[i] = q.has_seed ? q.seed_hi : 0;
      }
    } else if (seeded) {
      std::fill(lo.begin(), lo.end(), P.seed_lo);
      std::fill(hi.begin(), hi.end(), P.seed_hi);
    } else {
      std::random_device rd;  // rng::gen_seed (core/src/rng.rs:37-45): thread_rng
      for (size_t i = 0; i < N; ++i) lo[i] = ((uint64_t)rd() << 32) | rd();
    }
    if (!seeded) {  // GameConfig::seed_range (core/src/lib.rs:157-165), every env from its own config
      for (size_t i = 0; i < N; ++i) {
        const rg_params& q = per_env ? (*per_env)[i] : P;
        if (!q.has_seed_range) continue;
        if (q.seed_range_hi <= q.seed_range_lo) return fail(set_err(b, RG_ERR_SETTING, "empty seed_range"));
        lo[i] = q.seed_range_lo + lo[i] % (q.seed_range_hi - q.seed_range_lo);
      }
    }
    RG_TRY(cudaMemcpyAsync(b->d_u64, lo.data(), N * 8, cudaMemcpyHostToDevice, b->stream));
    RG_TRY(cudaMemcpyAsync(b->d_u64 + N, hi.data(), N * 8, cudaMemcpyHostToDevice, b->stream));
    RG_TRY(rg::launch_seed(d, b->d_u64, b->d_u64 + N, seeded ? 1 : 0, (int64_t)N, b->stream));
    RG_TRY(cudaStreamSynchronize(b->stream));
  }
  // GameStateImpl::new builds the game and takes the first PlayerState (state_impls.rs:20-37)
  RG_TRY(rg::launch_reset(d, b->stream));
  b->launches += 2;
  rc = rg_sync(b);
  if (rc != RG_OK) return fail(rc);
  rc = invalidate_prefetched(b);
  if (rc != RG_OK) return fail(rc);
#undef RG_TRY
  *out = b;
  return RG_OK;
}

// Empties every env's ring of prefetched games and queues all envs for the background generator
// (at creation, and whenever the seeds / episode counters the games were built for change).
int invalidate_prefetched(rg_batch* b) {
  if (!b->d.prefetch) return RG_OK;
  for (int i = 0; i < 2; ++i) RG_CUDA(b, cudaStreamSynchronize(b->bg[i]));
  const size_t N = (size_t)b->n;
  RG_CUDA(b, cudaMemsetAsync(b->d.sp_state, 0, N * rg::SP_DEPTH, b->stream));
  RG_CUDA(b, cudaMemsetAsync(b->d.sp_lock, 0, N * 4, b->stream));
  RG_CUDA(b, cudaMemsetAsync(b->d.sp_cancel, 0, N * rg::SP_DEPTH * 4, b->stream));
  std::vector<uint32_t> ids(N);
  for (size_t i = 0; i < N; ++i) ids[i] = (uint32_t)i;
  const uint32_t ctl[4] = {(uint32_t)N, 0u, 0u, 0u};  // tail = N, served window empty
  RG_CUDA(b, cudaMemcpyAsync(b->d.refill_ring, ids.data(), N * 4, cudaMemcpyHostToDevice, b->stream));
  RG_CUDA(b, cudaMemcpyAsync(b->d.refill_ctl, ctl, sizeof(ctl), cudaMemcpyHostToDevice, b->stream));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));  // ids / ctl are stack memory
  return RG_OK;
}

// Queue background pass k (window slot k % 8, fixed by k_step_end of the step just queued) after
// everything queued on the main stream so far. Passes alternate between two streams so that a pass
// that is still building does not delay the next one; the main stream waits for pass k-2 before it
// goes on, which bounds the time from a refill request to a ready game to about two steps.
int kick_prefetch(rg_batch* b) {
  if (!b->d.prefetch) return RG_OK;
  const int64_t k = b->passes++;
  const int i = (int)(k & 1);
  RG_CUDA(b, cudaEventRecord(b->ev_main, b->stream));
  RG_CUDA(b, cudaStreamWaitEvent(b->bg[i], b->ev_main, 0));
  RG_CUDA(b, rg::launch_prefetch(b->d, b->prefetch_warps, (int)(k % 8), b->bg[i]));
  if (k >= 2 && cudaEventQuery(b->ev_bg[i]) != cudaSuccess) {  // pass k-2 (same stream, already queued before pass k)
    (void)cudaGetLastError();
    RG_CUDA(b, cudaStreamWaitEvent(b->stream, b->ev_bg[i], 0));
  }
  RG_CUDA(b, cudaEventRecord(b->ev_bg[i], b->bg[i]));
  b->launches += 1;
  b->prefetch_running = true;
  return RG_OK;
}

// Queue the skeleton pass that serves the requests made up to the end of the step just queued (window slot
// step % 8, fixed by that step's step_end). Passes rotate over SPEC_STREAMS streams; the main stream waits for
// pass k - SPEC_STREAMS, so a pass may take that many steps without holding the steps up.
int kick_spec(rg_batch* b, int slot) {
  if (!b->d.spec) return RG_OK;
  const int64_t k = b->spec_passes++;
  const int i = (int)(k % rg_batch::SPEC_STREAMS);
  RG_CUDA(b, cudaEventRecord(b->ev_step, b->stream));
  RG_CUDA(b, cudaStreamWaitEvent(b->spec[i], b->ev_step, 0));
  RG_CUDA(b, rg::launch_spec_build(b->d, slot, b->spec[i]));
  // (a pass that old has finished long ago as a rule: ask the event on the host and queue a wait - one more operation
  // for the stream to retire at the end of a synced step - only if it has not)
  if (k >= rg_batch::SPEC_STREAMS && cudaEventQuery(b->ev_spec[i]) != cudaSuccess) {
    (void)cudaGetLastError();  // cudaErrorNotReady is an answer, not a failure
    RG_CUDA(b, cudaStreamWaitEvent(b->stream, b->ev_spec[i], 0));
  }
  RG_CUDA(b, cudaEventRecord(b->ev_spec[i], b->spec[i]));
  b->launches += 1;
  return RG_OK;
}

int copy_obs(rg_batch* b, rg_host_obs* out) {
  if (!out) return RG_OK;
  const DevBatch& d = b->d;
  const size_t N = (size_t)b->n;
  if (out->screen)
    RG_CUDA(b, cudaMemcpy2DAsync(out->screen, d.C, d.screen, d.CP, d.C, N, cudaMemcpyDeviceToHost, b->stream));
  if (out->history) {
    if (!b->d_hist_bytes) RG_CUDA(b, dev_alloc(b, &b->d_hist_bytes, N * d.C));
    RG_CUDA(b, rg::launch_unpack_hist(d, b->d_hist_bytes, b->stream));
    b->launches += 1;
    RG_CUDA(b, cudaMemcpyAsync(out->history, b->d_hist_bytes, N * d.C, cudaMemcpyDeviceToHost, b->stream));
  }
  if (out->status) RG_CUDA(b, cudaMemcpyAsync(out->status, d.status, N * 40, cudaMemcpyDeviceToHost, b->stream));
  if (out->reward) RG_CUDA(b, cudaMemcpyAsync(out->reward, d.reward, N * 4, cudaMemcpyDeviceToHost, b->stream));
  if (out->done) RG_CUDA(b, cudaMemcpyAsync(out->done, d.done, N, cudaMemcpyDeviceToHost, b->stream));
  if (out->message) RG_CUDA(b, cudaMemcpyAsync(out->message, d.message, N * 4, cudaMemcpyDeviceToHost, b->stream));
  if (out->error) RG_CUDA(b, cudaMemcpyAsync(out->error, d.error, N, cudaMemcpyDeviceToHost, b->stream));
  return RG_OK;
}

}  // namespace

extern "C" {

const char* rg_version(void) { return "rogue-gym_b200 0.1 (sm_100a)"; }

const char* rg_last_error(rg_batch* b) { return b ? b->err.c_str() : g_last_error.c_str(); }

int rg_create_from_params(const rg_params* p, int64_t n_envs, int64_t max_steps, int device, rg_batch** out) {
  if (!p) return set_err(nullptr, RG_ERR_ARG, "rg_create_from_params: params is null");
  return create_impl(*p, nullptr, n_envs, max_steps, device, out);
}

int rg_create(const char* const* cfg_json, int64_t n_cfg, int64_t n_envs, int64_t max_steps, int device, rg_batch** out) {
  if (!cfg_json || n_cfg < 1 || (n_cfg != 1 && n_cfg != n_envs))
    return set_err(nullptr, RG_ERR_ARG, "rg_create: n_cfg must be 1 or n_envs");
  char ebuf[512];
  rg_params P;
  int rc = rg_parse_config(cfg_json[0], &P, ebuf, sizeof(ebuf));
  if (rc != RG_OK) return set_err(nullptr, rc, ebuf);
  if (n_cfg == 1) return create_impl(P, nullptr, n_envs, max_steps, device, out);
  std::vector<rg_params> all((size_t)n_cfg);
  all[0] = P;
  for (int64_t i = 1; i < n_cfg; ++i) {
    rc = rg_parse_config(cfg_json[i], &all[(size_t)i], ebuf, sizeof(ebuf));
    if (rc != RG_OK) return set_err(nullptr, rc, ebuf);
    if (!same_geometry(P, all[(size_t)i]))
      return set_err(nullptr, RG_ERR_SETTING,
                     "Error in rogue-gym: Invalid Setting: the configs of one batch must agree on width, height, "
                     "room_num_x / room_num_y and the number of symbols (they fix the memory layout and the "
                     "observation shape); everything else may differ per env");
    char vbuf[256];
    rc = rg_validate_params(&all[(size_t)i], vbuf, sizeof(vbuf));
    if (rc != RG_OK) return set_err(nullptr, rc, vbuf);
  }
  return create_impl(P, &all, n_envs, max_steps, device, out);
}

void rg_destroy(rg_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (int i = 0; i < 2; ++i)
    if (b->bg[i]) cudaStreamSynchronize(b->bg[i]);
  if (b->side) cudaStreamSynchronize(b->side);
  if (b->mon) {
    cudaStreamSynchronize(b->mon);
    cudaStreamDestroy(b->mon);
  }
  if (b->ev_mon) cudaEventDestroy(b->ev_mon);
  if (b->ev_fast) cudaEventDestroy(b->ev_fast);
  for (int i = 0; i < rg_batch::SPEC_STREAMS; ++i) {
    if (b->spec[i]) {
      cudaStreamSynchronize(b->spec[i]);
      cudaStreamDestroy(b->spec[i]);
    }
    if (b->ev_spec[i]) cudaEventDestroy(b->ev_spec[i]);
  }
  if (b->ev_step) cudaEventDestroy(b->ev_step);
  for (int i = 0; i < 2; ++i) {
    if (b->graph[i]) cudaGraphExecDestroy(b->graph[i]);
    if (b->graph_m[i]) cudaGraphExecDestroy(b->graph_m[i]);
  }
  if (b->mir) cudaStreamSynchronize(b->mir);
  if (b->mir) cudaStreamDestroy(b->mir);
  if (b->ev_mir) cudaEventDestroy(b->ev_mir);
  if (b->ev_fork) cudaEventDestroy(b->ev_fork);
  if (b->ev_join) cudaEventDestroy(b->ev_join);
  if (b->side) cudaStreamDestroy(b->side);
  if (b->ev_main) cudaEventDestroy(b->ev_main);
  for (int i = 0; i < 2; ++i) {
    if (b->ev_bg[i]) cudaEventDestroy(b->ev_bg[i]);
    if (b->bg[i]) cudaStreamDestroy(b->bg[i]);
  }
  for (void* p : b->dev_allocs) cudaFree(p);
  if (b->m_host) cudaFreeHost(b->m_host);
  if (b->h_count) cudaFreeHost(b->h_count);
  if (b->h_errflag) cudaFreeHost(b->h_errflag);
  if (b->h_actions) cudaFreeHost(b->h_actions);
  if (b->h_error) cudaFreeHost(b->h_error);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int rg_seed_first(rg_batch* b, const uint64_t* seed_lo, const uint64_t* seed_hi, int64_t count) {
  if (!b || !seed_lo) return set_err(b, RG_ERR_ARG, "rg_seed: null argument");
  if (count < 0 || count > b->n) return set_err(b, RG_ERR_ARG, "rg_seed_first: count out of range");
  if (count == 0) return RG_OK;
  const size_t N = (size_t)b->n, K = (size_t)count;
  RG_CUDA(b, cudaSetDevice(b->device));
  {
    int rc = invalidate_prefetched(b);
    if (rc != RG_OK) return rc;
  }
  RG_CUDA(b, cudaMemcpyAsync(b->d_u64, seed_lo, K * 8, cudaMemcpyHostToDevice, b->stream));
  if (seed_hi) RG_CUDA(b, cudaMemcpyAsync(b->d_u64 + N, seed_hi, K * 8, cudaMemcpyHostToDevice, b->stream));
  RG_CUDA(b, rg::launch_seed(b->d, b->d_u64, seed_hi ? b->d_u64 + N : nullptr, 1, count, b->stream));
  b->launches += 1;
  RG_CUDA(b, cudaStreamSynchronize(b->stream));  // the host arrays may be pageable
  return RG_OK;
}

int rg_seed(rg_batch* b, const uint64_t* seed_lo, const uint64_t* seed_hi) {
  return rg_seed_first(b, seed_lo, seed_hi, b ? b->n : 0);
}

int rg_reset(rg_batch* b) {
  if (!b) return set_err(b, RG_ERR_ARG, "rg_reset: null batch");
  RG_CUDA(b, cudaSetDevice(b->device));
  int rc = invalidate_prefetched(b);
  if (rc != RG_OK) return rc;
  RG_CUDA(b, rg::launch_reset(b->d, b->stream));
  b->launches += 1;
  return RG_OK;
}

}  // extern "C"

namespace {
// One env-step on the batch's stream; with_mirror adds the host-mirror passes (rg_step_mirror).
int step_impl(rg_batch* b, const uint8_t* actions_dev, int auto_reset, bool with_mirror) {
  RG_CUDA(b, cudaSetDevice(b->device));
  auto_reset = auto_reset ? 1 : 0;
  const rg::MirrorArgs* mirror = with_mirror ? &b->margs : nullptr;
  cudaGraphExec_t* graphs = with_mirror ? b->graph_m : b->graph;
  const DevBatch& d = b->d;
  // the step kernels read the batch's own action buffer, so that the launch sequence has no
  // per-step argument and can be replayed as a graph
  if (actions_dev != b->d_actions)
    RG_CUDA(b, cudaMemcpyAsync(b->d_actions, actions_dev, (size_t)b->n, cudaMemcpyDeviceToDevice, b->stream));
  b->d.trace_step = (int32_t)(b->steps_launched++ % 512);
  if (b->use_graph) {
    if (!graphs[auto_reset]) {
      cudaGraph_t g = nullptr;
      RG_CUDA(b, cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
      // the host-facing step is kernels only: the scan reads the keys straight from the pinned staging buffer (mapped),
      // the last mirror pass writes the byte counter and the error flag into mapped host memory
      cudaError_t le = rg::launch_step(d, with_mirror ? b->h_actions_dev : b->d_actions, b->d_actions, auto_reset,
                                       b->step_streams(), mirror, b->sm_count);
      cudaError_t ce = cudaStreamEndCapture(b->stream, &g);
      if (le != cudaSuccess) return cuda_fail(b, le, "launch_step (capture)");
      if (ce != cudaSuccess) return cuda_fail(b, ce, "cudaStreamEndCapture");
      RG_CUDA(b, cudaGraphInstantiate(&graphs[auto_reset], g, 0));
      cudaGraphDestroy(g);
    }
    RG_CUDA(b, cudaGraphLaunch(graphs[auto_reset], b->stream));
  } else {
    RG_CUDA(b, rg::launch_step(d, with_mirror ? b->h_actions_dev : b->d_actions, b->d_actions, auto_reset, b->step_streams(),
                               mirror, b->sm_count));
  }
  b->launches += 8 + (with_mirror ? 2 : 0);
  {
    int rc = kick_spec(b, (int)((b->steps_launched - 1) % 8));
    if (rc != RG_OK) return rc;
  }
  if (auto_reset && (b->auto_steps++ % b->prefetch_every) == 0)
    return kick_prefetch(b);  // refill the next-episode buffers consumed so far
  return RG_OK;
}
}  // namespace

extern "C" {

int rg_step(rg_batch* b, const uint8_t* actions_dev, int auto_reset) {
  if (!b || !actions_dev) return set_err(b, RG_ERR_ARG, "rg_step: null argument");
  return step_impl(b, actions_dev, auto_reset, false);
}

int rg_stats(rg_batch* b, uint64_t* out8) {
  if (!b || !out8) return set_err(b, RG_ERR_ARG, "rg_stats: null argument");
  RG_CUDA(b, cudaSetDevice(b->device));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  for (int i = 0; i < 2; ++i)
    if (b->bg[i]) RG_CUDA(b, cudaStreamSynchronize(b->bg[i]));
  for (int i = 0; i < rg_batch::SPEC_STREAMS; ++i)
    if (b->spec[i]) RG_CUDA(b, cudaStreamSynchronize(b->spec[i]));
  RG_CUDA(b, cudaMemcpy(out8, b->d.stats, 64, cudaMemcpyDeviceToHost));
  return RG_OK;
}

int rg_trace(rg_batch* b, uint64_t* out, int64_t* steps_launched) {
  if (!b || !out) return set_err(b, RG_ERR_ARG, "rg_trace: null argument");
  if (!b->d.trace) return set_err(b, RG_ERR_ARG, "rg_trace: tracing is off (set RG_TRACE=1 before creating the batch)");
  RG_CUDA(b, cudaSetDevice(b->device));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  for (int i = 0; i < 2; ++i)
    if (b->bg[i]) RG_CUDA(b, cudaStreamSynchronize(b->bg[i]));
  RG_CUDA(b, cudaMemcpy(out, b->d.trace, 512 * rg::TRACE_KERNELS * 2 * 8, cudaMemcpyDeviceToHost));
  {  // read-and-clear: the slots are min/max accumulators and wrap every 512 steps
    std::vector<unsigned long long> init(512 * rg::TRACE_KERNELS * 2);
    for (size_t i = 0; i < init.size(); ++i) init[i] = (i & 1) ? 0ull : ~0ull;
    RG_CUDA(b, cudaMemcpy(b->d.trace, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
  }
  if (steps_launched) *steps_launched = b->steps_launched;
  return RG_OK;
}

int rg_quiesce(rg_batch* b) {
  if (!b) return set_err(b, RG_ERR_ARG, "rg_quiesce: null batch");
  if (b->d.prefetch && b->prefetch_running)
    for (int i = 0; i < 2; ++i) RG_CUDA(b, cudaStreamWaitEvent(b->stream, b->ev_bg[i], 0));
  if (b->d.spec && b->spec_passes > 0)
    for (int i = 0; i < rg_batch::SPEC_STREAMS && i < b->spec_passes; ++i)
      RG_CUDA(b, cudaStreamWaitEvent(b->stream, b->ev_spec[i], 0));
  return RG_OK;
}

int rg_sync(rg_batch* b) {
  if (!b) return set_err(b, RG_ERR_ARG, "rg_sync: null batch");
  RG_CUDA(b, cudaMemcpyAsync(b->h_errflag, b->d.errflag, sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  if (*b->h_errflag == 0) return RG_OK;
  // like the conductor, surface the first failing env's error (thread_impls.rs:65-68)
  RG_CUDA(b, cudaMemcpyAsync(b->h_error, b->d.error, (size_t)b->n, cudaMemcpyDeviceToHost, b->stream));
  RG_CUDA(b, cudaMemsetAsync(b->d.errflag, 0, sizeof(uint32_t), b->stream));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  for (int64_t i = 0; i < b->n; ++i)
    if (b->h_error[i]) {
      int code = b->h_error[i];
      return set_err(b, code, std::string("Error in rogue-gym: ") + status_text(code) + " (env " + std::to_string(i) + ")");
    }
  return RG_OK;
}

int rg_step_host(rg_batch* b, const uint8_t* actions_host, int auto_reset, rg_host_obs* out) {
  if (!b || !actions_host) return set_err(b, RG_ERR_ARG, "rg_step_host: null argument");
  RG_CUDA(b, cudaSetDevice(b->device));
  RG_CUDA(b, cudaMemcpyAsync(b->d_actions, actions_host, (size_t)b->n, cudaMemcpyHostToDevice, b->stream));
  int rc = rg_step(b, b->d_actions, auto_reset);
  if (rc != RG_OK) return rc;
  rc = copy_obs(b, out);
  if (rc != RG_OK) return rc;
  return rg_sync(b);
}

int rg_fetch(rg_batch* b, rg_host_obs* out) {
  if (!b) return set_err(b, RG_ERR_ARG, "rg_fetch: null batch");
  RG_CUDA(b, cudaSetDevice(b->device));
  int rc = copy_obs(b, out);
  if (rc != RG_OK) return rc;
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  return RG_OK;
}

int rg_fetch_terminal(rg_batch* b, uint8_t* out_host) {
  if (!b || !out_host) return set_err(b, RG_ERR_ARG, "rg_fetch_terminal: null argument");
  RG_CUDA(b, cudaSetDevice(b->device));
  uint8_t* tmp = reinterpret_cast<uint8_t*>(b->d_u64);  // [2N] u64 scratch
  RG_CUDA(b, rg::launch_state_terminal(b->d, tmp, b->stream));
  b->launches += 1;
  RG_CUDA(b, cudaMemcpyAsync(out_host, tmp, (size_t)b->n, cudaMemcpyDeviceToHost, b->stream));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  return RG_OK;
}

int rg_set_panic_policy(rg_batch* b, int policy) {
  if (!b || (policy != 0 && policy != 1)) return set_err(b, RG_ERR_ARG, "rg_set_panic_policy: policy must be 0 (sticky) or 1 (terminal)");
  if (b->d.panic_policy == policy) return RG_OK;
  RG_CUDA(b, cudaSetDevice(b->device));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  b->d.panic_policy = policy;
  for (int i = 0; i < 2; ++i) {  // the captured step graphs hold a copy of the batch descriptor: capture again
    if (b->graph[i]) cudaGraphExecDestroy(b->graph[i]);
    if (b->graph_m[i]) cudaGraphExecDestroy(b->graph_m[i]);
    b->graph[i] = b->graph_m[i] = nullptr;
  }
  return RG_OK;
}

int rg_views_get(rg_batch* b, rg_views* out) {
  if (!b || !out) return set_err(b, RG_ERR_ARG, "rg_views_get: null argument");
  out->n_envs = b->n;
  out->width = b->d.W;
  out->height = b->d.H;
  out->cell_stride = b->d.CP;
  out->hist_stride = b->d.HB;
  out->screen = b->d.screen;
  out->history_bits = b->d.hist;
  out->status = b->d.status;
  out->reward = b->d.reward;
  out->done = b->d.done;
  out->message = b->d.message;
  out->error = b->d.error;
  return RG_OK;
}

static int ensure_level_seen(rg_batch* b) {
  if (b->d_level_seen) return RG_OK;
  RG_CUDA(b, dev_alloc(b, &b->d_level_seen, (size_t)b->n));
  std::vector<int32_t> ones((size_t)b->n, 1);
  RG_CUDA(b, cudaMemcpyAsync(b->d_level_seen, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice, b->stream));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  return RG_OK;
}

int rg_train_reset(rg_batch* b) {
  if (!b) return set_err(b, RG_ERR_ARG, "rg_train_reset: null batch");
  RG_CUDA(b, cudaSetDevice(b->device));
  if (b->d_level_seen) {
    std::vector<int32_t> ones((size_t)b->n, 1);
    RG_CUDA(b, cudaMemcpyAsync(b->d_level_seen, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice, b->stream));
    RG_CUDA(b, cudaStreamSynchronize(b->stream));
  }
  return RG_OK;
}

int rg_step_train(rg_batch* b, const void* action_idx_dev, int index_bytes, int mode, uint32_t status_flag, int with_hist,
                  float* obs_out_dev, float* reward_out_dev, float stair_reward) {
  if (!b || !action_idx_dev) return set_err(b, RG_ERR_ARG, "rg_step_train: null argument");
  if (index_bytes != 0 && index_bytes != 1 && index_bytes != 4 && index_bytes != 8)
    return set_err(b, RG_ERR_ARG, "rg_step_train: index_bytes must be 0 (ASCII keys), 1, 4 or 8");
  RG_CUDA(b, cudaSetDevice(b->device));
  RG_CUDA(b, rg::launch_keys_from_index(b->d, action_idx_dev, index_bytes, b->d_actions, b->stream));
  b->launches += 1;
  int rc = step_impl(b, b->d_actions, 1, false);
  if (rc != RG_OK) return rc;
  if (obs_out_dev) {
    rc = rg_encode(b, mode, status_flag, with_hist, obs_out_dev, nullptr);
    if (rc != RG_OK) return rc;
  }
  if (reward_out_dev) {
    rc = ensure_level_seen(b);
    if (rc != RG_OK) return rc;
    RG_CUDA(b, rg::launch_train_reward(b->d, stair_reward, b->d_level_seen, reward_out_dev, b->stream));
    b->launches += 1;
  }
  return RG_OK;
}

int rg_mirror_get(rg_batch* b, rg_host_obs* out, uint8_t** history_bits) {
  if (!b || !out) return set_err(b, RG_ERR_ARG, "rg_mirror_get: null argument");
  RG_CUDA(b, cudaSetDevice(b->device));
  if (!b->m_host) {
    const DevBatch& d = b->d;
    const size_t N = (size_t)b->n;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_screen = 0, o_hist = o_screen + up(N * d.C), o_status = o_hist + up(N * d.HB),
                 o_reward = o_status + up(N * 40), o_message = o_reward + up(N * 4), o_done = o_message + up(N * 4),
                 o_error = o_done + up(N), total = o_error + up(N);
    void* hp = nullptr;
    RG_CUDA(b, cudaHostAlloc(&hp, total, cudaHostAllocMapped));
    memset(hp, 0, total);  // == the zeroed shadows below
    void* dp = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&dp, hp, 0);
    if (e == cudaSuccess && getenv("RG_MIRROR_NOHOST")) {  // diagnostic: the stores go to device memory (the mirror stays stale)
      void* scratch = nullptr;
      e = cudaMalloc(&scratch, total);
      if (e == cudaSuccess) {
        b->dev_allocs.push_back(scratch);
        dp = scratch;
      }
    }
    if (e != cudaSuccess) {
      cudaFreeHost(hp);
      return cuda_fail(b, e, "cudaHostGetDevicePointer");
    }
    auto fill = [&](rg_host_obs& o, uint8_t*& hist, void* base) {
      uint8_t* p = static_cast<uint8_t*>(base);
      o.screen = p + o_screen;
      o.history = nullptr;
      hist = p + o_hist;
      o.status = reinterpret_cast<uint32_t*>(p + o_status);
      o.reward = reinterpret_cast<int32_t*>(p + o_reward);
      o.message = reinterpret_cast<uint32_t*>(p + o_message);
      o.done = p + o_done;
      o.error = p + o_error;
    };
    fill(b->m_obs, b->m_hist_host, hp);
    fill(b->m_dev, b->m_hist_dev, dp);
    auto setup = [&]() -> cudaError_t {  // the shadows start zeroed, like the host block
      cudaError_t e2;
      if ((e2 = dev_alloc(b, &b->ms_screen, N * d.CP)) != cudaSuccess) return e2;
      if ((e2 = dev_alloc(b, &b->ms_hist, N * d.HB)) != cudaSuccess) return e2;
      const size_t flat = (N * 40 + 63) / 64 * 64 + 2 * ((N * 4 + 63) / 64 * 64) + 2 * ((N + 63) / 64 * 64);
      if ((e2 = dev_alloc(b, &b->ms_flat, flat)) != cudaSuccess) return e2;
      if ((e2 = dev_alloc(b, &b->m_count, 1)) != cudaSuccess) return e2;
      if (!b->h_count && (e2 = cudaHostAlloc(&b->h_count, 2 * sizeof(uint64_t), cudaHostAllocMapped)) != cudaSuccess) return e2;
      b->h_count[0] = b->h_count[1] = 0;
      if ((e2 = dev_alloc(b, &b->m_seq, 1)) != cudaSuccess) return e2;
      if ((e2 = cudaMemsetAsync(b->m_seq, 0, 8, b->stream)) != cudaSuccess) return e2;
      if (const char* v = getenv("RG_MIRROR_SPIN")) b->mirror_spin = v[0] != '0';
      if ((e2 = dev_alloc(b, &b->m_ticket, 1)) != cudaSuccess) return e2;
      if ((e2 = cudaMemsetAsync(b->m_ticket, 0, 4, b->stream)) != cudaSuccess) return e2;
      // k_mirror_lines on a few SMs (default) | k_mirror on every SM (RG_MIRROR_MODE=direct)
      b->margs.mode = 1;
      b->margs.confined_sms = 16;
      if (const char* v = getenv("RG_MIRROR_MODE")) b->margs.mode = !strcmp(v, "direct") ? 0 : 1;
      if (const char* v = getenv("RG_MIRROR_SMS")) b->margs.confined_sms = std::max(1, atoi(v));
      b->margs.fast_first = 1;
      if (const char* v = getenv("RG_MIRROR_FAST_FIRST")) b->margs.fast_first = v[0] != '0';
      if ((e2 = cudaMemsetAsync(b->ms_screen, 0, N * d.CP, b->stream)) != cudaSuccess) return e2;
      if ((e2 = cudaMemsetAsync(b->ms_hist, 0, N * d.HB, b->stream)) != cudaSuccess) return e2;
      if ((e2 = cudaMemsetAsync(b->ms_flat, 0, flat, b->stream)) != cudaSuccess) return e2;
      return cudaMemsetAsync(b->m_count, 0, 8, b->stream);
    };
    e = setup();
    if (e != cudaSuccess) {  // nothing half-built stays behind (the device pieces are freed with the batch)
      cudaFreeHost(hp);
      return cuda_fail(b, e, "rg_mirror_get: allocating the shadows");
    }
    b->m_host = hp;
    b->m_bytes = total;
    rg::MirrorArgs& m = b->margs;
    m.h_screen = b->m_dev.screen; m.h_hist = b->m_hist_dev; m.h_status = b->m_dev.status; m.h_reward = b->m_dev.reward;
    m.h_done = b->m_dev.done; m.h_message = b->m_dev.message; m.h_error = b->m_dev.error;
    m.s_screen = b->ms_screen; m.s_hist = b->ms_hist; m.s_flat = b->ms_flat; m.bytes = b->m_count;
    m.ticket = b->m_ticket;
    m.h_base = static_cast<uint8_t*>(dp);
    m.wide = 1;
    if (const char* e = getenv("RG_MIRROR_WIDE")) m.wide = e[0] != '0';
    {
      void *hb = nullptr, *he = nullptr;
      RG_CUDA(b, cudaHostGetDevicePointer(&hb, b->h_count, 0));
      RG_CUDA(b, cudaHostGetDevicePointer(&he, b->h_errflag, 0));
      m.h_bytes = static_cast<unsigned long long*>(hb);
      m.h_seq = m.h_bytes + 1;
      m.seq = b->m_seq;
      m.h_errflag = static_cast<uint32_t*>(he);
    }
    int rc = rg_mirror_sync(b, nullptr);  // the mirror starts out current
    if (rc != RG_OK && rc != RG_ERR_PANIC && rc != RG_ERR_INVALID_INPUT && rc != RG_ERR_IGNORED_INPUT) return rc;
  }
  if (history_bits && !b->margs.with_hist) {
    // the visited map is mirrored only once somebody asks for it (a third of the small PCIe writes of a step):
    // from now on it is; mark every row so that the next pass brings the host copy up to date
    RG_CUDA(b, cudaStreamSynchronize(b->stream));
    b->margs.with_hist = 1;
    for (int i = 0; i < 2; ++i) {  // the captured graphs hold a copy of the mirror arguments
      if (b->graph_m[i]) cudaGraphExecDestroy(b->graph_m[i]);
      b->graph_m[i] = nullptr;
    }
    RG_CUDA(b, cudaMemsetAsync(b->d.scr_rows, 0xFF, (size_t)b->n * 8, b->stream));
    int rc = rg_mirror_sync(b, nullptr);
    if (rc != RG_OK && rc != RG_ERR_PANIC && rc != RG_ERR_INVALID_INPUT && rc != RG_ERR_IGNORED_INPUT) return rc;
  }
  *out = b->m_obs;
  if (history_bits) *history_bits = b->m_hist_host;
  return RG_OK;
}

}  // extern "C"
namespace {
// Waits for the publishing mirror pass launched last. Its last block writes the pass counter into mapped host memory
// after everything else (system-scope fences in between), so the host can spin on that word: the mirror is readable
// ~5 us earlier than cudaStreamSynchronize reports the stream idle. A failed launch or a device fault never publishes:
// the stream is polled now and then.
int wait_mirror(rg_batch* b) {
  const uint64_t want = ++b->m_seq_expected;
  if (b->mirror_spin) {
    volatile uint64_t* seq = b->h_count + 1;
    for (uint32_t spins = 1; *seq < want; ++spins) {
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
      if ((spins & 0x3FFFu) == 0) {
        cudaError_t q = cudaStreamQuery(b->stream);
        if (q == cudaSuccess) break;  // drained: the word is there (or the pass never ran - caught below)
        if (q != cudaErrorNotReady) return cuda_fail(b, q, "host mirror pass");
      }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    if (*seq >= want) return RG_OK;
  }
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  return RG_OK;
}
}  // namespace
extern "C" {

int rg_mirror_sync(rg_batch* b, uint64_t* bytes_to_host) {
  if (!b) return set_err(b, RG_ERR_ARG, "rg_mirror_sync: null batch");
  if (!b->m_host) return set_err(b, RG_ERR_ARG, "rg_mirror_sync: call rg_mirror_get first");
  RG_CUDA(b, cudaSetDevice(b->device));
  RG_CUDA(b, rg::launch_mirror(b->d, b->margs, b->sm_count, b->stream));  // publishes the byte counter itself
  b->launches += 1;
  if (int rc = wait_mirror(b)) return rc;  // the mirror is readable now
  if (bytes_to_host) *bytes_to_host = *b->h_count;
  if (*b->h_errflag == 0) return RG_OK;
  const uint8_t* err = b->m_obs.error;
  for (int64_t i = 0; i < b->n; ++i)
    if (err[i]) {
      const int code = err[i];
      return set_err(b, code, std::string("Error in rogue-gym: ") + status_text(code) + " (env " + std::to_string(i) + ")");
    }
  return RG_OK;
}

int rg_step_mirror(rg_batch* b, const uint8_t* actions_host, int auto_reset, uint64_t* bytes_to_host) {
  if (!b || !actions_host) return set_err(b, RG_ERR_ARG, "rg_step_mirror: null argument");
  if (!b->m_host) return set_err(b, RG_ERR_ARG, "rg_step_mirror: call rg_mirror_get first");
  RG_CUDA(b, cudaSetDevice(b->device));
  // One graph launch, kernels only: the actions go through a pinned, mapped staging buffer (the caller's array may be
  // pageable and changes from call to call, a graph needs a fixed source) that k_step_scan reads over PCIe; the step
  // with its mirror passes - most envs are written back beside the other kernels, the rest after the step's last
  // kernel, which also writes the byte counter and the error flag into mapped host memory.
  memcpy(b->h_actions, actions_host, (size_t)b->n);
  int rc = step_impl(b, b->d_actions, auto_reset, true);
  if (rc != RG_OK) return rc;
  if (int rc2 = wait_mirror(b)) return rc2;  // the mirror is readable now
  if (bytes_to_host) *bytes_to_host = *b->h_count;
  if (*b->h_errflag == 0) return RG_OK;
  // Some env raised an error (a sticky reference-panic env does so on every step): report the first failing env like
  // rg_sync does (thread_impls.rs:65-68) - from the mirror's own error array, which is current; no further device work
  // (the publishing pass has cleared the device flag).
  const uint8_t* err = b->m_obs.error;
  for (int64_t i = 0; i < b->n; ++i)
    if (err[i]) {
      const int code = err[i];
      return set_err(b, code, std::string("Error in rogue-gym: ") + status_text(code) + " (env " + std::to_string(i) + ")");
    }
  return RG_OK;
}

void* rg_stream(rg_batch* b) { return b ? (void*)b->stream : nullptr; }
int64_t rg_launch_count(rg_batch* b) { return b ? b->launches : 0; }

int rg_encode_channels(const rg_batch* b, int mode, uint32_t status_flag, int with_hist) {
  if (!b || (mode != 0 && mode != 1)) return -1;
  int base = mode == 0 ? 1 : (int)b->P.symbols;
  return base + __builtin_popcount(status_flag & 0x1FFu) + (with_hist ? 1 : 0);
}

int rg_encode(rg_batch* b, int mode, uint32_t status_flag, int with_hist, float* out_dev, int* channels) {
  if (!b || !out_dev) return set_err(b, RG_ERR_ARG, "rg_encode: null argument");
  int ch = rg_encode_channels(b, mode, status_flag, with_hist);
  if (ch < 0) return set_err(b, RG_ERR_ARG, "rg_encode: mode must be 0 (gray) or 1 (symbol)");
  if (channels) *channels = ch;
  RG_CUDA(b, rg::launch_encode(b->d, mode, status_flag & 0x1FFu, with_hist, ch, out_dev, b->stream));
  b->launches += 1;
  return RG_OK;
}

int rg_encode_compact(rg_batch* b, uint8_t* sym_out_dev, int32_t* status_out_dev, uint8_t* hist_out_dev) {
  if (!b || !sym_out_dev) return set_err(b, RG_ERR_ARG, "rg_encode_compact: null argument");
  RG_CUDA(b, cudaSetDevice(b->device));
  RG_CUDA(b, rg::launch_encode_compact(b->d, sym_out_dev, status_out_dev, hist_out_dev, b->sm_count, b->stream));
  b->launches += 1;
  return RG_OK;
}

int rg_state_hash(rg_batch* b, uint64_t* out_host) {
  if (!b || !out_host) return set_err(b, RG_ERR_ARG, "rg_state_hash: null argument");
  RG_CUDA(b, cudaSetDevice(b->device));
  RG_CUDA(b, rg::launch_state_hash(b->d, b->d_u64, b->stream));
  b->launches += 1;
  RG_CUDA(b, cudaMemcpyAsync(out_host, b->d_u64, (size_t)b->n * 8, cudaMemcpyDeviceToHost, b->stream));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  return RG_OK;
}

int rg_export_floors(rg_batch* b, uint8_t* surface_dev, uint8_t* attr_dev, int16_t* rooms_dev) {
  if (!b) return set_err(b, RG_ERR_ARG, "rg_export_floors: null batch");
  RG_CUDA(b, cudaSetDevice(b->device));
  const DevBatch& d = b->d;
  const size_t N = (size_t)b->n;
  if (surface_dev)
    RG_CUDA(b, cudaMemcpy2DAsync(surface_dev, d.C, d.surface, d.CP, d.C, N, cudaMemcpyDeviceToDevice, b->stream));
  if (attr_dev) RG_CUDA(b, cudaMemcpy2DAsync(attr_dev, d.C, d.attr, d.CP, d.C, N, cudaMemcpyDeviceToDevice, b->stream));
  if (rooms_dev) {
    RG_CUDA(b, rg::launch_export_rooms(d, rooms_dev, b->stream));
    b->launches += 1;
  }
  return RG_OK;
}

int rg_test_move_enemy(rg_batch* b, int64_t env, int fx, int fy, int tx, int ty, int* kind, int* nx, int* ny) {
  if (!b || env < 0 || env >= b->n) return set_err(b, RG_ERR_ARG, "rg_test_move_enemy: bad env");
  if (fx < 0 || fy < 0 || fx >= b->d.W || fy >= b->d.H || tx < 0 || ty < 0 || tx >= b->d.W || ty >= b->d.H)
    return set_err(b, RG_ERR_ARG, "rg_test_move_enemy: coordinate out of range");
  RG_CUDA(b, cudaSetDevice(b->device));
  RG_CUDA(b, rg::launch_test_move_enemy(b->d, env, fx, fy, tx, ty, b->d_out3, b->stream));
  b->launches += 1;
  int h[3];
  RG_CUDA(b, cudaMemcpyAsync(h, b->d_out3, sizeof(h), cudaMemcpyDeviceToHost, b->stream));
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  if (kind) *kind = h[0];
  if (nx) *nx = h[1];
  if (ny) *ny = h[2];
  return RG_OK;
}

int rg_dump_env(rg_batch* b, int64_t env, rg_dump* out) {
  if (!b || !out || env < 0 || env >= b->n) return set_err(b, RG_ERR_ARG, "rg_dump_env: bad argument");
  const DevBatch& d = b->d;
  RG_CUDA(b, cudaSetDevice(b->device));
  if (out->cache_maps) {  // DistCache maps are evaluated lazily on the device: finish this env's maps first
    RG_CUDA(b, rg::launch_complete_maps(d, env, env + 1, b->stream));
    b->launches += 1;
  }
  RG_CUDA(b, cudaStreamSynchronize(b->stream));
  EnvState st;
  RG_CUDA(b, cudaMemcpy(&st, d.st + env, sizeof(st), cudaMemcpyDeviceToHost));
  rg_dump_scalars& s = out->s;
  memset(&s, 0, sizeof(s));
  s.level = st.level;
  s.px = st.px;
  s.py = st.py;
  s.hp = st.hp;
  s.hp_max = st.hp_max;
  s.exp = st.exp;
  s.plevel = st.plevel;
  s.food_left = st.food_left;
  s.quiet = st.quiet;
  s.gold = st.gold;
  s.ui_dead = st.ui_dead;
  s.steps = (int32_t)st.steps;
  s.is_terminal = st.is_terminal;
  s.message = st.message;
  s.error = st.error;
  s.n_cache = st.cache_n;
  memcpy(s.status, st.status, sizeof(s.status));
  memcpy(s.rng, st.rng, sizeof(s.rng));
  const int nrooms = b->P.room_num_x * b->P.room_num_y;
  struct M { int x, y, kind, hp, active, level, defense, exp; };
  std::vector<M> mons;
  const int la = (int)((uint32_t)st.level > b->P.amulet_level ? (uint32_t)st.level - b->P.amulet_level : 0u);
  for (int i = 0; i < nrooms; ++i) {
    const rg::MonD& m = st.mon[i];
    if (!(m.flags & rg::MF_PRESENT)) continue;
    const rg_enemy_kind& k = b->P.enemies[m.kind];
    mons.push_back(M{m.x, m.y, m.kind, m.hp, (m.flags & rg::MF_ACTIVE) ? 1 : 0, k.level + la, k.defense - la, (int)m.exp});
  }
  std::sort(mons.begin(), mons.end(), [](const M& a, const M& c) { return a.x != c.x ? a.x < c.x : a.y < c.y; });
  s.n_monsters = (int32_t)mons.size();
  if (out->monsters) {
    memset(out->monsters, 0, sizeof(int32_t) * RG_MAX_ROOMS * 8);
    for (size_t i = 0; i < mons.size(); ++i) memcpy(out->monsters + i * 8, &mons[i], sizeof(M));
  }
  std::vector<std::pair<int, uint32_t>> items;
  for (int i = 0; i < nrooms; ++i)
    if (st.item_pos[i] != 0xFFFF) items.push_back({st.item_pos[i], st.item_amt[i]});
  std::sort(items.begin(), items.end());
  s.n_items = (int32_t)items.size();
  if (out->items) {
    memset(out->items, 0, sizeof(int32_t) * RG_MAX_ROOMS * 3);
    for (size_t i = 0; i < items.size(); ++i) {
      out->items[i * 3 + 0] = items[i].first % d.W;
      out->items[i * 3 + 1] = items[i].first / d.W;
      out->items[i * 3 + 2] = (int32_t)items[i].second;
    }
  }
  if (out->surface) RG_CUDA(b, cudaMemcpy(out->surface, d.surface + env * d.CP, d.C, cudaMemcpyDeviceToHost));
  if (out->attr) RG_CUDA(b, cudaMemcpy(out->attr, d.attr + env * d.CP, d.C, cudaMemcpyDeviceToHost));
  if (out->cache_xy)
    for (int i = 0; i < RG_DIST_CACHE; ++i) {
      bool live = i < st.cache_n;
      int slot = (st.cache_head + i) % RG_DIST_CACHE;
      out->cache_xy[i * 2] = live ? st.cache_x[slot] : -1;
      out->cache_xy[i * 2 + 1] = live ? st.cache_y[slot] : -1;
      if (live && out->cache_maps)
        RG_CUDA(b, cudaMemcpy(out->cache_maps + (size_t)i * d.C, d.dist + ((size_t)env * RG_DIST_CACHE + slot) * d.CP,
                              (size_t)d.C * 2, cudaMemcpyDeviceToHost));
    }
  if (out->rooms) {
    memset(out->rooms, 0, sizeof(int32_t) * RG_MAX_ROOMS * 8);
    for (int i = 0; i < nrooms; ++i) {
      const rg::RoomD& r = st.rooms[i];
      int32_t* o = out->rooms + i * 8;
      o[0] = r.kind;
      o[1] = (r.flags & rg::RF_DARK) ? 1 : 0;
      o[2] = (r.flags & rg::RF_VISITED) ? 1 : 0;
      o[3] = (r.flags & rg::RF_GOLD) ? 1 : 0;
      o[4] = r.x0;
      o[5] = r.y0;
      o[6] = r.x1;
      o[7] = r.y1;
    }
  }
  return RG_OK;
}

// PlayerState::{gray,symbol}_image[_with_hist] for detached PlayerState values (python/src/lib.rs:158-205):
// the states are uploaded, encoded by the same kernel as rg_encode, and the images copied back.
int rg_encode_states(rg_batch* b, int64_t n, const uint8_t* screens, const uint8_t* history, const uint32_t* status,
                     int mode, uint32_t status_flag, int with_hist, float* out_host, int* channels) {
  if (!b || n < 1 || !screens || !status || !out_host || (with_hist && !history))
    return set_err(b, RG_ERR_ARG, "rg_encode_states: bad argument");
  int ch = rg_encode_channels(b, mode, status_flag, with_hist);
  if (ch < 0) return set_err(b, RG_ERR_ARG, "rg_encode_states: mode must be 0 (gray) or 1 (symbol)");
  if (channels) *channels = ch;
  RG_CUDA(b, cudaSetDevice(b->device));
  DevBatch t = b->d;
  t.n = n;
  const size_t N = (size_t)n;
  uint8_t *scr = nullptr, *hist = nullptr, *err = nullptr;
  uint32_t* st = nullptr;
  float* out = nullptr;
  std::vector<uint8_t> packed(N * t.HB, 0);
  if (history)
    for (size_t e = 0; e < N; ++e)
      for (int i = 0; i < t.C; ++i)
        if (history[e * t.C + i]) packed[e * t.HB + (i >> 3)] |= (uint8_t)(1u << (i & 7));
  int rc = RG_OK;
  auto run = [&]() -> cudaError_t {
    cudaError_t e;
    if ((e = cudaMalloc(&scr, N * t.CP)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&hist, N * t.HB)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&err, N)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&st, N * 40)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&out, N * (size_t)ch * t.C * sizeof(float))) != cudaSuccess) return e;
    if ((e = cudaMemcpy2DAsync(scr, t.CP, screens, t.C, t.C, N, cudaMemcpyHostToDevice, b->stream)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(hist, packed.data(), N * t.HB, cudaMemcpyHostToDevice, b->stream)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(st, status, N * 40, cudaMemcpyHostToDevice, b->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(err, 0, N, b->stream)) != cudaSuccess) return e;
    t.screen = scr;
    t.hist = hist;
    t.status = st;
    t.error = err;
    if ((e = rg::launch_encode(t, mode, status_flag & 0x1FFu, with_hist, ch, out, b->stream)) != cudaSuccess) return e;
    b->launches += 1;
    if ((e = cudaMemcpyAsync(out_host, out, N * (size_t)ch * t.C * sizeof(float), cudaMemcpyDeviceToHost, b->stream)) !=
        cudaSuccess)
      return e;
    std::vector<uint8_t> herr(N);
    if ((e = cudaMemcpyAsync(herr.data(), err, N, cudaMemcpyDeviceToHost, b->stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(b->stream)) != cudaSuccess) return e;
    for (size_t i = 0; i < N; ++i)
      if (herr[i]) {
        rc = set_err(b, RG_ERR_SETTING, "Error in rogue-gym: Invalid tile in the screen of state " + std::to_string(i) +
                                            ", while max is " + std::to_string(b->P.symbols - 1));
        // this launch raised the shared flag; it is not a step error
        cudaMemsetAsync(b->d.errflag, 0, sizeof(uint32_t), b->stream);
        cudaStreamSynchronize(b->stream);
        break;
      }
    return cudaSuccess;
  };
  cudaError_t ce = run();
  cudaFree(scr);
  cudaFree(hist);
  cudaFree(err);
  cudaFree(st);
  cudaFree(out);
  if (ce != cudaSuccess) return cuda_fail(b, ce, "rg_encode_states");
  return rc;
}

}  // extern "C"
