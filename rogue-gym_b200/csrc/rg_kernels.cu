// rg_kernels.cu — __global__ entry points (warp per env) and their launchers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "rg_device.cuh"
#include "rg_launch.h"

namespace rg {

#ifndef RG_WPB
#define RG_WPB 1
#endif
#ifndef RG_HOT_MIN_BLOCKS
#define RG_HOT_MIN_BLOCKS 32
#endif
// One warp per block: a block's registers and shared memory are released the moment its env is
// done, so a warp that runs a BFS does not pin three finished neighbours (measured: 4 warps/block
// 0.668 ms per step, 1 warp/block 0.626 ms at 64 registers, 65 536 envs).
constexpr int WARPS_PER_BLOCK = RG_WPB;
// The generator kernels are the opposite case: ~15 k instructions of divergent code run by a few
// hundred warps beside the step kernels. Spread one warp per SM they evict the step kernels' code from
// every SM's instruction caches (measured: the player kernel takes 216 us instead of 140 us while a
// background pass runs); packed into a few many-warp blocks they disturb only the SMs they sit on.
// k_prefetch takes its warps per block from the launch (DevBatch::pf_wpb, at most PF_MAX_WPB and as many
// as fit in shared memory). The full-path kernel handles ~10 envs per step: one warp per block, spread.
constexpr int PF_MAX_WPB = 16;
constexpr int SPEC_WPB = 8;  // warps per block of k_spec_build (a few dozen requests per step: a handful of blocks)
constexpr int PF_EXCLUSIVE_SMEM = 227 * 1024 - 4608;  // leaves less than one step-kernel block's worth (4.4 KB + 1 KB reserved)
#ifndef RG_GEN_WPB
#define RG_GEN_WPB 1
#endif
constexpr int GEN_WPB = RG_GEN_WPB;  // k_step_gen (218 registers)
RG_DEV size_t warp_smem(const DevBatch& b) { return 2 * (size_t)b.CP + sizeof(EnvState) + 16; }

// ---- staging through shared memory with the bulk-copy engine (TMA, 1-D): one elected lane
// issues cp.async.bulk for the env's EnvState and tile planes, all in flight at once, completion
// on an mbarrier; write-back is a bulk shared->global group. (A loop of per-lane 128-bit loads
// left the kernels waiting on one DRAM round trip per iteration: 48 % of the player kernel's
// stall samples were long-scoreboard in load_grid / fill_ctx.)
RG_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Stager {
  uint32_t bar;    // shared address of this warp's mbarrier
  uint32_t phase;  // parity of the next completion
};
RG_DEV Stager stager_init(const DevBatch& b, unsigned char* base) {
  Stager s;
  s.bar = smem_u32(base + 2 * (size_t)b.CP + sizeof(EnvState));
  s.phase = 0;
  if ((threadIdx.x & 31) == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s.bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  return s;
}
RG_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
RG_DEV void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
enum { PL_NONE = 0, PL_S = 1, PL_A = 2, PL_BOTH = 3 };
// Starts the bulk loads of EnvState (+ the requested planes) of `env` into the warp's region.
RG_DEV void stage_issue(const DevBatch& b, Stager& s, unsigned char* base, int64_t env, int planes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses to this region first
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    const uint32_t bytes = (uint32_t)sizeof(EnvState) + ((planes & PL_S) ? b.CP : 0) + ((planes & PL_A) ? b.CP : 0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s.bar), "r"(bytes) : "memory");
    bulk_g2s(smem_u32(base + 2 * (size_t)b.CP), b.st + env, (uint32_t)sizeof(EnvState), s.bar);
    if (planes & PL_S) bulk_g2s(smem_u32(base), b.surface + env * b.CP, (uint32_t)b.CP, s.bar);
    if (planes & PL_A) bulk_g2s(smem_u32(base + b.CP), b.attr + env * b.CP, (uint32_t)b.CP, s.bar);
  }
}
// Waits for the loads started by stage_issue.
RG_DEV void stage_wait(Stager& s) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(s.bar), "r"(s.phase)
        : "memory");
  } while (!ok);
  s.phase ^= 1u;
}
RG_DEV void stage_load(const DevBatch& b, Stager& s, unsigned char* base, int64_t env, int planes) {
  stage_issue(b, s, base, env, planes);
  stage_wait(s);
}

// Stage one env into shared memory and fill the context.
RG_DEV void fill_ctx(const DevBatch& b, Ctx& c, Stager& sg, unsigned char* base, int64_t env, int planes,
                     bool already_staged = false) {
  if (!already_staged) stage_load(b, sg, base, env, planes);
  c.soff = (uint32_t)(base - rg_smem);
  c.S = base;
  c.A = base + b.CP;
  c.st = reinterpret_cast<EnvState*>(base + 2 * (size_t)b.CP);
  c.col_room = b.room_lut;
  c.row_room = b.room_lut + 160;
  c.P = b.cfg_idx ? b.P + b.cfg_idx[env] : b.P;
  c.W = b.W; c.H = b.H; c.C = b.C; c.CP = b.CP; c.WW = b.WW;
  c.lane = threadIdx.x & 31;
  c.nx = b.nx; c.ny = b.ny;
  c.rsx = b.rsx; c.rsy = b.rsy;
  c.nrooms = c.nx * c.ny;
  c.g_screen = b.screen + env * b.CP;
  c.g_hist = b.hist + env * b.HB;
  c.g_rows = b.scr_rows + env;
  c.g_walk = b.walk + env * (int64_t)(b.H * b.WW);
  c.g_dist = b.dist + env * (int64_t)NCACHE * b.CP;
  c.g_bfs = b.bfs + env * (int64_t)NCACHE * 2 * b.H * b.WW;
  c.g_wsnap = b.wsnap + env * (int64_t)NCACHE * b.H * b.WW;
  c.g_spec_seq = nullptr;
  c.g_spec_hits = nullptr;
  if (b.spec) {
    c.g_spec_S = b.spec_S + env * b.CP;
    c.g_spec_A = b.spec_A + env * b.CP;
    c.g_spec_rooms = b.spec_rooms + env * MAX_ROOMS;
    c.g_spec_tag = b.spec_tag + env;
    c.g_spec_seq = b.spec_seq + env;
    c.g_spec_hits = b.stats + RGS_SPEC_HITS;
  }
  c.redraw = c.status_upd = c.dead = c.msg = c.hist_done = c.a_dirty = c.s_dirty = c.panic = 0;
  c.rd.load(c.st->rng);
  c.ri.load(c.st->rng + 4);
  c.re.load(c.st->rng + 8);
}
// Write-back: EnvState always, a plane only when it was modified. One bulk group; the issuing
// lane waits until shared memory has been read (the region may be reused or the block may exit).
RG_DEV void write_back(const DevBatch& b, Ctx& c, int64_t env, bool with_s, bool with_a, EnvState* st_dst, uint8_t* s_dst,
                       uint8_t* a_dst) {
  __syncwarp();
  {  // what k_step_fast needs to know about the monsters, recomputed wherever the state is written back
    uint32_t present = 0, active = 0;
    if (c.lane < MAX_ROOMS) {
      const MonD m = c.st->mon[c.lane];
      present = (m.flags & MF_PRESENT) ? 1u : 0u;
      active = (present && (m.flags & MF_ACTIVE)) ? 1u : 0u;
      c.st->mon_xy[c.lane] = (uint16_t)(m.x | (m.y << 8));
    }
    const uint32_t pm = __ballot_sync(RG_FULL, present), am = __ballot_sync(RG_FULL, active);
    if (c.lane == 0) {
      c.st->mon_present = (uint16_t)pm;
      c.st->mon_active = (uint16_t)am;
      c.rd.store(c.st->rng);
      c.ri.store(c.st->rng + 4);
      c.re.store(c.st->rng + 8);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (c.lane == 0) {
    bulk_s2g(st_dst + env, smem_u32(c.st), (uint32_t)sizeof(EnvState));
    if (with_s) bulk_s2g(s_dst + env * b.CP, smem_u32(c.S), (uint32_t)b.CP);
    if (with_a) bulk_s2g(a_dst + env * b.CP, smem_u32(c.A), (uint32_t)b.CP);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
}
RG_DEV void close_env(const DevBatch& b, Ctx& c, int64_t env) {
  write_back(b, c, env, c.s_dirty != 0, c.a_dirty != 0, b.st, b.surface, b.attr);
}
// `done` is what the step RETURNS: after an auto-reset the conductor flags the copy it hands back while the
// worker's own state - the fresh game - is not terminal (thread_impls.rs:69-79); -1 = the state's own flag
RG_DEV void emit_obs(const DevBatch& b, Ctx& c, int64_t env, int32_t reward, uint8_t err, int done = -1) {
  if (c.lane < 10) b.status[env * 10 + c.lane] = c.st->status[c.lane];
  if (c.lane == 10) b.reward[env] = reward;
  if (c.lane == 11) b.done[env] = done < 0 ? c.st->is_terminal : (uint8_t)done;
  if (c.lane == 12) b.message[env] = c.st->message;
  if (c.lane == 13) {
    b.error[env] = err;
    if (err) atomicOr(b.errflag, 1u << err);
  }
}

// Optional timeline (RG_TRACE=1): first start / last end of every kernel of a step over all its blocks,
// read back with rg_trace.
enum { TK_PLAYER = 0, TK_MONSTERS = 1, TK_FINISH = 2, TK_FULL = 3, TK_RESETS = 4, TK_PREFETCH = 5, TK_PLAYER_B = 6,
       TK_MONSTERS_B = 7, TK_MIRROR1 = 8, TK_MIRROR2 = 9, TK_SCAN = 10 };
RG_DEV unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct TraceScope {
  unsigned long long* slot;
  RG_DEV TraceScope(const DevBatch& b, int kernel) {
    slot = nullptr;
    // every block reports (tracing is a diagnostic mode): the last block to finish is what the next kernel waits for
    if (b.trace && threadIdx.x == 0) {
      slot = b.trace + ((size_t)(kernel == TK_PREFETCH ? b.trace_step : (int)(*b.dstep % 512u)) * TRACE_KERNELS + kernel) * 2;
      atomicMin(slot, gtime());
    }
  }
  RG_DEV ~TraceScope() {
    if (slot) atomicMax(slot + 1, gtime());
  }
};

// Work the hot step kernel hands to the generation kernel (one entry per env, any order).
enum : uint32_t { DEFER_STEP = 0u, DEFER_RESET = 0x80000000u };
RG_DEV void count_event(const DevBatch& b, const Ctx& c, int which) {
  if (c.lane == 0) atomicAdd(b.stats + which, 1ull);
}
RG_DEV void defer(const DevBatch& b, Ctx& c, int64_t env, uint32_t code, int parity) {
  if (c.lane == 0) {
    atomicAdd(b.stats + (code ? RGS_SYNC_RESET : RGS_FULL_STEP), 1ull);
    // two lists: full-path steps come from k_step_scan (their kernel runs beside the player and monster
    // kernels on a side stream), synchronous resets are known only when finish_env has run
    if (code) b.full_path[env] = FP_RESET;  // finished by the reset pass: not yet final for the host mirror
    uint32_t* cnt = (code ? b.reset_count : b.defer_count) + parity;
    uint32_t* list = code ? b.reset_list : b.defer_list;
    list[atomicAdd(cnt, 1u)] = (uint32_t)env | code;
  }
}

// Speculative descents: once per level (and again after the dungeon stream was drawn from during play), when the
// player comes within SPEC_RADIUS of the stair, ask the background builder for the next level's skeleton.
RG_DEV void push_spec_request(const DevBatch& b, int64_t env) {
  const uint32_t t = atomicAdd(b.spec_ctl, 1u);
  b.spec_ring[t & (b.spec_cap - 1u)] = (uint32_t)env;
}
RG_DEV bool near_stair(int W, int px, int py, uint32_t stair_pos) {
  if (stair_pos == 0xFFFFu) return false;
  const int sy = (int)stair_pos / W, sx = (int)stair_pos - sy * W;
  return max(abs(px - sx), abs(py - sy)) <= SPEC_RADIUS;
}
// for the warp kernels: `st` is the env's state staged in shared memory, about to be written back as the live state
RG_DEV void maybe_request_spec(const DevBatch& b, Ctx& c, int64_t env) {
  if (!b.spec) return;
  EnvState* st = c.st;
  if (st->spec_req || st->error || st->ui_dead) return;
  if (!near_stair(c.W, st->px, st->py, st->stair_pos)) return;
  __syncwarp();
  st->spec_req = 1;
  if (c.lane == 0) push_spec_request(b, env);
  __syncwarp();
}

// ThreadWorker::run Instruction::Reset for every env (python/src/thread_impls.rs:117-124)
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, RG_HOT_MIN_BLOCKS) k_reset(DevBatch b) {
  unsigned char* const smem = rg_smem;
  Ctx c;
  const int warp = threadIdx.x >> 5;
  const int64_t env = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
  if (env >= b.n) return;
  unsigned char* const base = smem + (size_t)warp * warp_smem(b);
  Stager sg = stager_init(b, base);
  fill_ctx(b, c, sg, base, env, PL_NONE);
  reset_env<true>(c);
  uint8_t err = 0;
  if (c.panic) {
    c.st->error = RG_ERR_PANIC;
    err = RG_ERR_PANIC;
  }
  compose(c);
  emit_obs(b, c, env, 0, err);
  maybe_request_spec(b, c, env);
  close_env(b, c, env);
}

RG_DEV void store_state(const DevBatch& b, Ctx& c, int64_t env) {
  write_back(b, c, env, false, false, b.st, b.surface, b.attr);
}
RG_DEV bool has_active_monster(const Ctx& c) {
  bool any = false;
#pragma unroll 1
  for (int m = 0; m < c.nrooms; ++m) {
    const uint8_t f = c.st->mon[m].flags;
    any = any || ((f & MF_PRESENT) && (f & MF_ACTIVE));
  }
  return any;
}

// ---------------------------------------------------------------------------------------------
// One env-step = GameStateImpl::react (python/src/state_impls.rs:51-79) and, with auto_reset,
// the conductor's "reset terminal envs and return the fresh state flagged terminal"
// (thread_impls.rs:69-79). It runs as a short pipeline of kernels, each small enough to stay
// resident in the instruction caches (the single-kernel version spent most of its stall samples
// on instruction fetch: 17-30 k SASS instructions executed divergently by independent warps):
//
//   k_step_scan     one thread per env, key + two state words: descents and MoveUntil go to the full-path list, envs
//                   with an active monster to the first slow list. Three branches run side by side from here.
//   k_step_fast     (main stream) one THREAD per remaining env: finishes the cheap common steps itself (see there);
//                   what needs warp-wide work goes to the second slow list. full_path[] says who ends the step.
//   k_step_gen      full-path list, high-priority side stream, beside the two kernels below: the whole
//                   step compiled as one piece with the floor generator. A second, normally empty pass
//                   after the monster kernel does the reset half of terminal steps that found no
//                   prefetched game.
//   k_step_player   a slow list (first list: own stream, beside k_step_fast; second list: after it), one warp per
//                   env: key -> action -> player move / attack / pickup / search, hunger,
//                   heal. Envs with an active monster go to the monster list; every other env is
//                   finished here (finish_env).
//   k_step_monsters monster list only: coin flips, lazy-BFS chase, attacks, then finish_env.
//   finish_env      message / status / step count / terminal, compose, observation; a terminal env
//                   under auto_reset takes its prefetched next game or goes to the reset list.
//
// The hand-over between the phases is EnvState::f_flags / f_msg / f_gold_before.
// ---------------------------------------------------------------------------------------------
// Tell the background generator that env's ring of prefetched games has a free slot.
RG_DEV void request_refill(const DevBatch& b, const Ctx& c, int64_t env) {
  if (b.prefetch && c.lane == 0) {
    const uint32_t t = atomicAdd(b.refill_ctl, 1u);
    b.refill_ring[t & (b.refill_cap - 1u)] = (uint32_t)env;
  }
}

// Copies env's prefetched game (planes, screen, history, walk rows, EnvState) over the live one.
RG_DEV void swap_in_prefetched(const DevBatch& b, Ctx& c, int64_t env, int64_t sp) {
  const int n16 = b.CP / 16;
  const uint4* s0 = reinterpret_cast<const uint4*>(b.sp_surface + sp * b.CP);
  const uint4* s1 = reinterpret_cast<const uint4*>(b.sp_attr + sp * b.CP);
  const uint4* s2 = reinterpret_cast<const uint4*>(b.sp_screen + sp * b.CP);
  uint4* d0 = reinterpret_cast<uint4*>(b.surface + env * b.CP);
  uint4* d1 = reinterpret_cast<uint4*>(b.attr + env * b.CP);
  uint4* d2 = reinterpret_cast<uint4*>(b.screen + env * b.CP);
  for (int i = c.lane; i < n16; i += 32) {
    d0[i] = s0[i];
    d1[i] = s1[i];
    d2[i] = s2[i];
  }
  const uint4* h0 = reinterpret_cast<const uint4*>(b.sp_hist + sp * b.HB);
  uint4* h1 = reinterpret_cast<uint4*>(b.hist + env * b.HB);
  for (int i = c.lane; i < b.HB / 16; i += 32) h1[i] = h0[i];
  const uint32_t* w0 = b.sp_walk + sp * (int64_t)(b.H * b.WW);
  uint32_t* w1 = b.walk + env * (int64_t)(b.H * b.WW);
  for (int i = c.lane; i < b.H * b.WW; i += 32) w1[i] = w0[i];
  if (c.lane == 0) b.scr_rows[env] = ~0ull;
  const uint4* e0 = reinterpret_cast<const uint4*>(b.sp_st + sp);
  uint4* e1 = reinterpret_cast<uint4*>(c.st);
  __syncwarp();
  for (int i = c.lane; i < (int)(sizeof(EnvState) / 16); i += 32) e1[i] = e0[i];
  __syncwarp();
  c.rd.load(c.st->rng);
  c.ri.load(c.st->rng + 4);
  c.re.load(c.st->rng + 8);
}

// Last part of a step for an env whose state is staged in shared memory (python/src/state_impls.rs:
// 56-77): message flags, displayed status, step count, terminal; under auto_reset the next episode
// is moved in (prefetched) or requested (k_step_gen); compose if a Redraw was emitted; observation.
// Runs at the end of k_step_player (envs without an active monster) or of k_step_monsters.
RG_DEV void finish_env(const DevBatch& b, Ctx& c, int64_t env, int auto_reset, int parity) {
  EnvState* st = c.st;
  const uint32_t gold_before = st->f_gold_before;
  uint8_t err = 0;
  // Panic policy. A state in which the reference panics kills that env's worker thread and with it the whole
  // conductor (thread_impls.rs:111-135). "sticky" (default): the env is frozen and reports RG_ERR_PANIC on every
  // step. "terminal" (rg_set_panic_policy / RG_PANIC_POLICY=terminal, auto-reset steps only): the step that hits
  // the state reports done = 1 with error RG_ERR_PANIC once and the env goes on with a fresh episode.
  const bool revive = c.panic && b.panic_policy == 1 && auto_reset;
  if (c.panic && !revive) {
    st->error = RG_ERR_PANIC;
    err = RG_ERR_PANIC;
  } else {
    if (revive) {
      st->is_terminal = 1;
      err = RG_ERR_PANIC;
    } else {
      st->message = c.msg;
      if (c.status_upd) refresh_status(c);
      st->steps += 1;
      st->is_terminal = (c.dead || (int64_t)st->steps >= b.max_steps) ? 1 : 0;
    }
    if (st->is_terminal && auto_reset) {
      // the game for episode e+1 lives in ring slot (e+1) % SP_DEPTH
      const int64_t sp = env * SP_DEPTH + (int64_t)((st->episode + 1) % SP_DEPTH);
      bool ready = b.prefetch && *reinterpret_cast<volatile uint8_t*>(b.sp_state + sp) == 1;
      if (ready) {
        __threadfence();
        // built for exactly this episode? (a synchronous reset may have overtaken the background pass)
        if (*reinterpret_cast<volatile uint32_t*>(&b.sp_st[sp].episode) != st->episode + 1) {
          ready = false;
          count_event(b, c, RGS_PREFETCH_STALE);
          __syncwarp();
          if (c.lane == 0) *reinterpret_cast<volatile uint8_t*>(b.sp_state + sp) = 0;
        }
      }
      if (ready) {
        // The next episode of this env was generated ahead of time (k_prefetch): move it in.
        count_event(b, c, RGS_SWAP_IN);
        swap_in_prefetched(b, c, env, sp);
        const uint8_t perr = st->error;  // the fresh game's own (generation) panic, if any
        const int32_t d0 = (int32_t)st->status[1] - (int32_t)gold_before;
        emit_obs(b, c, env, d0 > 0 ? d0 : 0, err ? err : perr, 1);
        maybe_request_spec(b, c, env);
        store_state(b, c, env);
        __threadfence();
        __syncwarp();
        if (c.lane == 0) *reinterpret_cast<volatile uint8_t*>(b.sp_state + sp) = 0;
        __threadfence();
        request_refill(b, c, env);
        return;
      }
      // not ready (or prefetch off): the fresh game is built synchronously by k_step_gen; a background
      // build of the same episode that is still under way must not be published afterwards
      if (b.prefetch && c.lane == 0) *reinterpret_cast<volatile uint32_t*>(b.sp_cancel + sp) = st->episode + 1;
      request_refill(b, c, env);
      if (c.lane == 0) {  // hand-over to the reset pass
        b.reward[env] = (int32_t)gold_before;
        b.error[env] = err;
      }
      store_state(b, c, env);
      defer(b, c, env, DEFER_RESET, parity);
      return;
    }
    if (c.redraw) compose(c);
  }
  const int32_t diff = (int32_t)st->status[1] - (int32_t)gold_before;
  emit_obs(b, c, env, diff > 0 ? diff : 0, err);
  maybe_request_spec(b, c, env);
  close_env(b, c, env);
}

// ---------------------------------------------------------------------------------------------
// k_step_fast: first kernel of a step, ONE THREAD PER ENV. It classifies every env and finishes the
// common cheap steps itself; whatever needs warp-wide work goes to the warp-per-env kernels through lists.
//   CL_FULL  descents and MoveUntil -> defer_list (k_step_gen, beside the other kernels)
//   CL_FAST  the whole step is done here: early-outs (sticky errors, steps > max_steps, invalid key, input after
//            death), NoOp, NoDownStair, a blocked move, a search with nothing hidden around, and a plain move -
//            no active monster, no door under either foot, no monster or gold on the target, no monster whose
//            drawing could change - for which only the <= 4x4 cells around the two positions are read, updated
//            (Cell::left / Cell::approached, field.rs:20-34) and redrawn
//   CL_SLOW  everything else -> slow_list (k_step_player)
// Why the partial redraw is the full redraw: the screen in HBM is persistent and, whenever dirty_rows == 0 and
// no monster is active, equal to a full compose of the state (every change of the planes, of the player's
// position or of a drawn entity is followed by a compose in the same step; only ACTIVE monsters move without
// one). A plain move changes the attr plane inside the two 3x3 neighbourhoods only, moves '@', and - by the
// eligibility rules - leaves every monster's drawn / hidden status as it was.
// Instruction count is the point: the warp-per-env kernel spends ~1 000 warp instructions per env-step with
// 31 of 32 lanes duplicating scalar work (ncu, round 1); here a warp retires 32 envs at once.
// ---------------------------------------------------------------------------------------------
enum { CL_FAST = 0, CL_FULL = 1, CL_SLOW = 2, CL_FAST_MOVE = 3, CL_FAST_LIGHT = 4 };

RG_DEV uint32_t load4(const uint8_t* plane, int idx) {  // the 4 bytes at plane[idx .. idx+3], idx >= 0, any alignment
  const uint32_t* w = reinterpret_cast<const uint32_t*>(plane + (idx & ~3));
  return __funnelshift_r(w[0], w[1], 8 * (idx & 3));
}
RG_DEV int cheb(int ax, int ay, int bx, int by) { return max(abs(ax - bx), abs(ay - by)); }
RG_DEV int lut_room(const DevBatch& b, int x, int y) {  // room_of without a Ctx
  if (x < 0 || y < 0 || x >= b.W || y >= b.H) return -1;
  const uint32_t cx = __ldg(b.room_lut + x), ry = __ldg(b.room_lut + 160 + y);
  return (cx == 0xFFu || ry == 0xFFu) ? -1 : (int)(ry * b.nx + cx);
}

RG_DEV uint32_t sel4(const uint32_t (&v)[4], int r) { return r == 0 ? v[0] : r == 1 ? v[1] : r == 2 ? v[2] : v[3]; }

// Loads are issued in two dependent rounds (the kernel is latency-bound: 2 048 warps, each thread a chain of
// scattered 16-byte reads): round 1 = the key and every piece of the hot block the action can need; round 2 = what
// the position decides - the <= 4 rows of both planes around the move, the visited-map byte, the sector table
// entries. Only the rare reads (a room record, the displayed gold) come later.
RG_DEV int fast_env(const DevBatch& b, int64_t env, uint8_t key, int auto_reset) {
  EnvState* const stp = b.st + env;
  const uint4* hot = reinterpret_cast<const uint4*>(stp);
  int d;
  const int act = map_key(key, d);
  const uint4 h0 = hot[0], h1 = hot[1], h2 = hot[2], h3 = hot[3], h4 = hot[4];
  uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0, i0 = m0, i1 = m0, h11 = m0;
  if (act == 0 && b.fast) {  // a move may need the monster and item tables and the stair: same round trip
    m0 = hot[5]; m1 = hot[6]; i0 = hot[7]; i1 = hot[8]; h11 = hot[11];
  }
  const int px = (int)(int16_t)(h0.x & 0xFFFFu), py = (int)(int16_t)(h0.x >> 16);
  const uint32_t is_terminal = h0.y & 0xFFu, ui_dead = (h0.y >> 8) & 0xFFu, serr = (h0.y >> 16) & 0xFFu;
  const uint32_t steps = h0.z;
  uint32_t food_left = h1.x, quiet = h1.y;
  const uint32_t gold = h1.z;
  int32_t hp = (int32_t)h2.y;
  const int32_t hp_max = (int32_t)h2.z, plevel = (int32_t)h2.w;
  const uint64_t dirty_rows = ((uint64_t)h3.y << 32) | h3.x;
  const uint32_t mon_present = h4.x & 0xFFFFu, mon_active = h4.x >> 16;
  const int W = b.W, H = b.H;

  // ---- the early-outs of GameStateImpl::react / react_to_key (state_impls.rs:52-54, core/src/lib.rs:314,322-327)
  {
    int early = -1;
    if (serr == RG_ERR_PANIC || serr == RG_ERR_SETTING) early = (int)serr;  // the reference's worker is gone
    else if ((int64_t)steps > b.max_steps) early = 0;
    else if (act < 0) early = RG_ERR_INVALID_INPUT;
    else if (ui_dead) early = RG_ERR_IGNORED_INPUT;
    if (early >= 0) {
      if (!b.fast || (early == RG_ERR_PANIC && b.panic_policy == 1 && auto_reset)) return CL_SLOW;
      b.reward[env] = 0;
      b.done[env] = (uint8_t)is_terminal;
      b.message[env] = h0.w;
      b.error[env] = (uint8_t)early;
      if (early) atomicOr(b.errflag, 1u << early);
      return CL_FAST;
    }
  }
  // (the full path - MoveUntil, DownStair on a stair - and the envs with an active monster are on k_step_scan's lists)
  if (!b.fast) return CL_SLOW;
  if ((int64_t)steps + 1 >= b.max_steps) return CL_SLOW;  // the step ends the episode: finish_env's business
  if (act != 4 && (mon_active != 0 || plevel >= 8)) return CL_SLOW;  // monster phase / heal draws on the enemy stream

  uint32_t msg = 0;
  bool moved = false;
  int nx = px, ny = py;
  uint64_t rows_touched = 0;
  const uint8_t* S = b.surface + env * b.CP;
  uint8_t* A = b.attr + env * b.CP;
  if (act == 3) {
    msg = MSG_NO_DOWNSTAIR;
  } else if (act == 2) {  // Floor::search floor.rs:349-370 with nothing to find: no draw, no change, an empty redraw
    if (dirty_rows != 0 || px < 1 || py < 1 || px + 1 >= W || py + 1 >= H) return CL_SLOW;
    uint32_t any = 0;
#pragma unroll
    for (int r = -1; r <= 1; ++r) any |= load4(A, (py + r) * W + px - 1) & 0x00FFFFFFu;
    if (any & (0x010101u * (A_HIDDEN | A_LOCKED))) return CL_SLOW;
  } else if (act == 0) {  // actions::move_player actions.rs:168-195
    nx = px + ddx(d);
    ny = py + ddy(d);
    bool can = nx >= 0 && ny >= 0 && nx < W && ny < H;
    if (can) {
      // the box that holds both 3x3 neighbourhoods, clipped to the field: rows y0..y1 (<= 4), columns x0..x1 (<= 4)
      const int x0 = max(min(px, nx) - 1, 0), x1 = min(max(px, nx) + 1, W - 1);
      const int y0 = max(min(py, ny) - 1, 0), y1 = min(max(py, ny) + 1, H - 1);
      const int pi = py * W + px, ni = ny * W + nx;
      // ---- round 2: every load whose address the position decides
      uint32_t sw[4], aw[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int rb = min(y0 + r, y1) * W + x0;  // rows past y1 repeat the last one (never used)
        sw[r] = load4(S, rb);
        aw[r] = load4(A, rb);
      }
      uint8_t* const hb = b.hist + env * b.HB + (ni >> 3);
      const uint32_t hist_byte = *hb;
      const uint64_t scr_rows = b.scr_rows[env];
      const int room_old = lut_room(b, px, py), room_new = lut_room(b, nx, ny);
      const uint32_t mons[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      int mon_room[MAX_ROOMS];
#pragma unroll
      for (int m = 0; m < MAX_ROOMS; ++m) {
        mon_room[m] = -1;
        if ((mon_present >> m) & 1u) {
          const uint32_t xy = (mons[m >> 1] >> (16 * (m & 1))) & 0xFFFFu;
          mon_room[m] = lut_room(b, (int)(xy & 0xFFu), (int)(xy >> 8));
        }
      }
      // ---- Floor::can_move_impl floor.rs:169-182, from the rows just loaded
      const int ro = py - y0, rn = ny - y0, co = px - x0, cn = nx - x0;
      const uint32_t s_n = (sel4(sw, rn) >> (8 * cn)) & 0xFFu, a_n = (sel4(aw, rn) >> (8 * cn)) & 0xFFu;
      const uint32_t a_p = (sel4(aw, ro) >> (8 * co)) & 0xFFu;
      can = can_walk((uint8_t)s_n) && !(a_n & (A_HIDDEN | A_LOCKED));
      if (can && is_diag(d))
        can = can_walk((uint8_t)((sel4(sw, ro) >> (8 * cn)) & 0xFFu)) && can_walk((uint8_t)((sel4(sw, rn) >> (8 * co)) & 0xFFu));
      if (can) {
        if (dirty_rows != 0) return CL_SLOW;
        if ((a_p | a_n) & A_DOOR) return CL_SLOW;  // leaves_room / enters_room
        const uint32_t items[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
        const uint32_t base0 = (uint32_t)(y0 * W + x0);
        uint32_t item_cells = 0;  // bit r*4+k: an item lies on box cell (y0+r, x0+k)
#pragma unroll
        for (int i = 0; i < MAX_ROOMS; ++i) {
          const uint32_t ip = (items[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
          if (ip == 0xFFFFu) continue;
          if (ip == (uint32_t)ni) return CL_SLOW;  // pickup: get_item actions.rs:206-231
          const uint32_t dd = ip - base0;
          if (dd <= (uint32_t)(3 * W + 3)) {
            const uint32_t r = dd >= (uint32_t)(3 * W) ? 3u : dd >= (uint32_t)(2 * W) ? 2u : dd >= (uint32_t)W ? 1u : 0u;
            const uint32_t k = dd - r * (uint32_t)W;
            if (k < 4u) item_cells |= 1u << (r * 4u + k);
          }
        }
        if (mon_present) {
#pragma unroll
          for (int m = 0; m < MAX_ROOMS; ++m) {
            if (!((mon_present >> m) & 1u)) continue;
            const uint32_t xy = (mons[m >> 1] >> (16 * (m & 1))) & 0xFFFFu;
            const int mx = (int)(xy & 0xFFu), my = (int)(xy >> 8);
            // on the target (player_attack), or close enough for its adjacency / its cell's attributes to change
            if (cheb(mx, my, px, py) <= 1 || cheb(mx, my, nx, ny) <= 1) return CL_SLOW;
            const int rm = mon_room[m];  // Dungeon::draw_enemy rogue/mod.rs:398-404 -> Floor::in_same_room
            if (rm >= 0 && (rm == room_old || rm == room_new)) {
              const RoomD r = stp->rooms[rm];
              const bool in_m = in_rect(r, mx, my);
              const bool so = rm == room_old && (r.kind == K_EMPTY || in_rect(r, px, py) == in_m);
              const bool sn = rm == room_new && (r.kind == K_EMPTY || in_rect(r, nx, ny) == in_m);
              if (so != sn) return CL_SLOW;
            }
          }
        }
        // ---- committed: Floor::player_out (floor.rs:298-312), Floor::player_in (:264-295), redraw of the box
        uint8_t* scr = b.screen + env * b.CP;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int y = y0 + r;
          if (y > y1) break;
          const int rb = y * W + x0;
          const bool drawn_row = y >= 1 && y < H - 1;  // rows 0 and H-1 are never written (python/src/lib.rs:44)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int x = x0 + k;
            if (x > x1) break;
            const uint32_t sv = (sw[r] >> (8 * k)) & 0xFFu, a_in = (aw[r] >> (8 * k)) & 0xFFu;
            uint32_t a = a_in;
            const bool near_old = abs(x - px) <= 1 && abs(y - py) <= 1, near_new = abs(x - nx) <= 1 && abs(y - ny) <= 1;
            if (!near_old && !near_new) continue;  // a corner of the box: nothing changes there (a monster may be drawn on it)
            if (near_old && sv == S_FLOOR && (a & A_DARK)) a &= ~(uint32_t)A_VISIBLE;  // Cell::left
            if (x == nx && y == ny) a |= A_VISITED;
            if (near_new) {  // Cell::approached
              const bool diag = x != nx && y != ny;
              if (!(diag && sv == S_PASSAGE) && !(a & A_HIDDEN)) a |= A_DRAWN | A_VISIBLE;
            }
            if (a != a_in) A[rb + k] = (uint8_t)a;
            if (drawn_row) {
              uint32_t ch = (a & A_VISIBLE) ? surface_tile((uint8_t)(sv & 7u)) : ' ';
              if (((item_cells >> (r * 4 + k)) & 1u) && (a & (A_VISIBLE | A_DRAWN))) ch = '*';
              if (x == nx && y == ny && (a & (A_VISIBLE | A_DRAWN))) ch = '@';
              scr[rb + k] = (uint8_t)ch;
            }
          }
          rows_touched |= 1ull << y;
        }
        *hb = (uint8_t)(hist_byte | (1u << (ni & 7)));  // Floor::history_map: the one cell that became visited
        b.scr_rows[env] = scr_rows | rows_touched;
        moved = true;
        if (b.spec) {
          const uint32_t tailw = h11.w;  // cache_head, spec_req, stair_pos
          if (!((tailw >> 8) & 0xFFu) && near_stair(W, nx, ny, tailw >> 16)) {
            stp->spec_req = 1;
            push_spec_request(b, env);
          }
        }
      }
    }
  }
  // ---- the turn: after_turn (actions.rs:67-80) = Player::turn_passed + heal (player.rs:163-176,221-240); no monster is active
  bool status_upd = false;
  if (act != 4) {
    const rg_params* P = b.cfg_idx ? b.P + b.cfg_idx[env] : b.P;
    food_left -= 1;
    if (food_left != 0) {
      const uint32_t hunger = P->hunger_time / 10;
      const bool hungry = food_left == hunger || food_left == hunger * 2;
      quiet += 1;
      const int amount = max(min((int)quiet + (plevel << 1) - 20, 1), 0);  // plevel < 8
      bool healed = false;
      if (amount > 0) {
        hp = min(hp + amount, hp_max);
        quiet = 0;
        healed = true;
      }
      status_upd = hungry || healed;
    }
    uint4* hw = reinterpret_cast<uint4*>(stp);
    hw[1] = make_uint4(food_left, quiet, gold, h1.w);
    if (hp != (int32_t)h2.y) hw[2] = make_uint4(h2.x, (uint32_t)hp, h2.z, h2.w);
    if (moved) {
      const uint64_t ov = (((uint64_t)h3.w << 32) | h3.z) | rows_touched;  // a superset is allowed (see EnvState::ov_rows)
      hw[3] = make_uint4(0u, 0u, (uint32_t)ov, (uint32_t)(ov >> 32));
    }
    int32_t reward = 0;
    if (status_upd) {  // RunTime::player_status -> Status::to_vec (refresh_status)
      const uint32_t hunger = P->hunger_time / 10;
      const uint32_t old_gold = stp->status[1];
      const uint32_t v[10] = {h2.x, gold, (uint32_t)hp, (uint32_t)hp_max, 16u, 16u, 0u, (uint32_t)plevel, h1.w,
                              food_left <= hunger ? 2u : (food_left <= hunger * 2 ? 1u : 0u)};
      uint2* sp = reinterpret_cast<uint2*>(stp->status);
      uint2* op = reinterpret_cast<uint2*>(b.status + env * 10);
#pragma unroll
      for (int i = 0; i < 5; ++i) sp[i] = op[i] = make_uint2(v[2 * i], v[2 * i + 1]);
      const int32_t diff = (int32_t)gold - (int32_t)old_gold;
      reward = diff > 0 ? diff : 0;
    }
    b.reward[env] = reward;
  } else {
    b.reward[env] = 0;
  }
  // ---- finish (state_impls.rs:56-77): message, step count; not terminal (checked above)
  {
    const uint32_t pos = moved ? ((uint32_t)(uint16_t)nx | ((uint32_t)(uint16_t)ny << 16)) : h0.x;
    reinterpret_cast<uint4*>(stp)[0] = make_uint4(pos, h0.y & 0xFFFFFF00u, steps + 1, msg);
  }
  b.done[env] = 0;
  b.message[env] = msg;
  b.error[env] = 0;
  return CL_FAST;
}

// First kernel of a step, one thread per env, two state words and the key: the envs whose step is known to need the
// warp kernels whatever else happens - the full path (descents, MoveUntil) and the envs with an active monster
// (player phase, then monster phase) - are listed at once, so that both chains start ~6 us into the step and run
// beside k_step_fast instead of after it. Everything else is left to k_step_fast (full_path[] = FP_FAST so far).
__global__ void __launch_bounds__(256) k_step_scan(DevBatch b, const uint8_t* __restrict__ actions, uint8_t* __restrict__ actions_out) {
  const int parity = (int)(*b.dstep & 1u);
  TraceScope trace(b, TK_SCAN);
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // next step's counters
    b.defer_count[parity ^ 1] = 0;
    b.reset_count[parity ^ 1] = 0;
    b.fast_count[(parity ^ 1) * 2] = b.fast_count[(parity ^ 1) * 2 + 1] = 0;
    for (int w = 0; w < 2; ++w) {
      b.slow_count[w * 2 + (parity ^ 1)] = 0;
      b.slow_count[4 + w * 2 + (parity ^ 1)] = 0;  // the player kernels' work cursors
      b.mon_count[w * 2 + (parity ^ 1)] = 0;
      b.mon_count[4 + w * 2 + (parity ^ 1)] = 0;   // the monster kernels' work cursors
    }
  }
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int cls = -1;
  if (env < b.n) {
    const uint4* hot = reinterpret_cast<const uint4*>(b.st + env);
    const uint4 h0 = hot[0];
    const uint32_t mon_active = hot[4].x >> 16;
    const uint32_t ui_dead = (h0.y >> 8) & 0xFFu, serr = (h0.y >> 16) & 0xFFu;
    int d;
    const uint8_t key = actions[env];
    if (actions_out != actions) actions_out[env] = key;  // the host-facing step reads the keys from mapped host memory once
    const int act = map_key(key, d);
    cls = CL_FAST;
    const bool early = serr == RG_ERR_PANIC || serr == RG_ERR_SETTING || (int64_t)h0.z > b.max_steps || act < 0 || ui_dead;
    if (!early) {
      const int px = (int)(int16_t)(h0.x & 0xFFFFu), py = (int)(int16_t)(h0.x >> 16);
      if (act == 1 || (act == 3 && b.surface[env * b.CP + py * b.W + px] == S_STAIR)) cls = CL_FULL;
      else if (act != 4 && mon_active != 0) cls = CL_SLOW;
    }
    b.full_path[env] = cls == CL_FAST ? FP_FAST : cls == CL_FULL ? FP_FULL : FP_PLAYER;
    // k_step_fast takes its envs grouped by what they do, so that the threads of a warp run the same code: the moves
    // (long: cells, monsters, items, redraw) in one list, everything else (early-outs, NoOp, NoDownStair, search) in
    // the other. Unsorted, a warp executed every path with 5 of its 32 threads active on average (ncu).
    if (cls == CL_FAST) cls = (!early && act == 0) ? CL_FAST_MOVE : CL_FAST_LIGHT;
  }
  const uint32_t slowM = __ballot_sync(RG_FULL, cls == CL_SLOW), fullM = __ballot_sync(RG_FULL, cls == CL_FULL);
  const uint32_t moveM = __ballot_sync(RG_FULL, cls == CL_FAST_MOVE), lightM = __ballot_sync(RG_FULL, cls == CL_FAST_LIGHT);
  uint32_t sbase = 0, fbase = 0, mbase = 0, lbase = 0;
  if (lane == 0) {
    if (slowM) sbase = atomicAdd(b.slow_count + parity, (uint32_t)__popc(slowM));
    if (fullM) {
      fbase = atomicAdd(b.defer_count + parity, (uint32_t)__popc(fullM));
      atomicAdd(b.stats + RGS_FULL_STEP, (unsigned long long)__popc(fullM));
    }
    if (moveM) mbase = atomicAdd(b.fast_count + parity * 2, (uint32_t)__popc(moveM));
    if (lightM) lbase = atomicAdd(b.fast_count + parity * 2 + 1, (uint32_t)__popc(lightM));
  }
  sbase = __shfl_sync(RG_FULL, sbase, 0);
  fbase = __shfl_sync(RG_FULL, fbase, 0);
  mbase = __shfl_sync(RG_FULL, mbase, 0);
  lbase = __shfl_sync(RG_FULL, lbase, 0);
  const uint32_t below = (1u << lane) - 1u;
  if (cls == CL_SLOW) b.slow_list[sbase + __popc(slowM & below)] = (uint32_t)env;
  if (cls == CL_FULL) b.defer_list[fbase + __popc(fullM & below)] = (uint32_t)env;
  if (cls == CL_FAST_MOVE) b.fast_list_m[mbase + __popc(moveM & below)] = (uint32_t)env;
  if (cls == CL_FAST_LIGHT) b.fast_list_l[lbase + __popc(lightM & below)] = (uint32_t)env;
}

__global__ void __launch_bounds__(128) k_step_fast(DevBatch b, const uint8_t* __restrict__ actions, int auto_reset) {
  const int parity = (int)(*b.dstep & 1u);
  TraceScope trace(b, TK_FINISH);
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t n_move = b.fast_count[parity * 2], n_light = b.fast_count[parity * 2 + 1];
  if (blockIdx.x == 0 && threadIdx.x == 0) b.fast_count[4] = n_move;  // (the mirror pass may outlive the step's parity)
  int cls = -1;
  int64_t env = -1;
  if (t < (int64_t)n_move) env = (int64_t)b.fast_list_m[t];
  else if (t < (int64_t)n_move + (int64_t)n_light) env = (int64_t)b.fast_list_l[t - (int64_t)n_move];
  if (env >= 0) {
    cls = fast_env(b, env, actions[env], auto_reset);
    if (cls != CL_FAST) b.full_path[env] = FP_PLAYER;
  }
  // warp-aggregated append of the leftovers to the second slow list
  const uint32_t slowM = __ballot_sync(RG_FULL, cls == CL_SLOW), fastM = __ballot_sync(RG_FULL, cls == CL_FAST);
  uint32_t sbase = 0;
  if (lane == 0) {
    if (slowM) sbase = atomicAdd(b.slow_count + 2 + parity, (uint32_t)__popc(slowM));
    if (fastM) atomicAdd(b.stats + RGS_FAST_STEPS, (unsigned long long)__popc(fastM));
  }
  sbase = __shfl_sync(RG_FULL, sbase, 0);
  if (cls == CL_SLOW) b.slow_list_b[sbase + __popc(slowM & ((1u << lane) - 1u))] = (uint32_t)env;
}

RG_DEV void player_env(const DevBatch& b, Stager& sg, unsigned char* base, int64_t env, const uint8_t* __restrict__ actions,
                       int auto_reset, int parity, int which) {
  // The env's state and both planes are requested first; the key is read while the bulk loads fly.
  stage_issue(b, sg, base, env, PL_BOTH);
  const uint8_t key = actions[env];
  stage_wait(sg);
  Ctx c;
  int d = 0;
  const int act = map_key(key, d);
  fill_ctx(b, c, sg, base, env, PL_BOTH, true);
  EnvState* st = c.st;
  // what is left to do for this env: 1 report `err` only, 2 the normal path
  int todo = 1;
  uint8_t err = 0;
  if (st->error == RG_ERR_PANIC && b.panic_policy == 1 && auto_reset) {
    // a game that was born in a panic state (or froze before the policy was switched): ends now, a fresh one follows
    st->f_gold_before = st->status[1];
    st->error = 0;
    c.panic = 1;
    finish_env(b, c, env, auto_reset, parity);
    return;
  }
  if (st->error == RG_ERR_PANIC || st->error == RG_ERR_SETTING) err = st->error;  // the reference's worker is gone
  else if ((int64_t)st->steps > b.max_steps) err = 0;                           // state_impls.rs:52-54
  else if (act < 0) err = RG_ERR_INVALID_INPUT;  // ErrorKind::InvalidInput: nothing changes (core/src/lib.rs:322-327)
  else if (st->ui_dead) err = RG_ERR_IGNORED_INPUT;  // UiState::Mordal(Grave) + Act => IgnoredInput (core/src/lib.rs:314)
  else todo = 2;
  if (todo == 1) {
    emit_obs(b, c, env, 0, err);
    return;
  }
  st->f_gold_before = st->status[1];
  if (act == 0 || act == 2 || act == 3) process_action<true>(c, act, d);  // act 3 here = NoDownStair: a turn passes
  if (act != 4 && !c.panic && has_active_monster(c)) {
    // hand over to the monster kernel: it runs the monster phase and then finishes the step
    st->f_msg = c.msg;
    st->f_flags = (uint8_t)((c.redraw ? SF_REDRAW : 0) | (c.status_upd ? SF_STATUS : 0));
    if (c.lane == 0) {
      (which ? b.mon_list_b : b.mon_list)[atomicAdd(b.mon_count + which * 2 + parity, 1u)] = (uint32_t)env;
      b.full_path[env] = FP_MONSTERS;
    }
    count_event(b, c, RGS_MONSTER_ENVS);
    close_env(b, c, env);
  } else {
    finish_env(b, c, env, auto_reset, parity);  // no monster moves this turn: the step ends here
  }
}

// One-warp blocks over a slow list: `which` = 0 the envs k_step_scan found with an active monster (this launch and
// its monster kernel run on a stream of their own beside k_step_fast), 1 the envs k_step_fast left over. Envs are
// handed out one at a time: their cost varies (a room reveal, a fight, an episode end with its 10 KB swap-in).
__global__ void __launch_bounds__(32, RG_HOT_MIN_BLOCKS)
k_step_player(DevBatch b, const uint8_t* __restrict__ actions, int auto_reset, int which) {
  unsigned char* const smem = rg_smem;
  const int parity = (int)(*b.dstep & 1u);
  TraceScope trace(b, which ? TK_PLAYER_B : TK_PLAYER);
  const uint32_t count = b.slow_count[which * 2 + parity];
  const uint32_t* const list = which ? b.slow_list_b : b.slow_list;
  Stager sg = stager_init(b, smem);
  for (;;) {
    uint32_t i = 0;
    if (threadIdx.x == 0) i = atomicAdd(b.slow_count + 4 + which * 2 + parity, 1u);
    i = __shfl_sync(RG_FULL, i, 0);
    if (i >= count) break;
    player_env(b, sg, smem, (int64_t)list[i], actions, auto_reset, parity, which);
    __syncwarp();
  }
}

// actions::move_active_enemies (actions.rs:82-119) for the envs that have an active monster
#ifndef RG_MON_MIN_BLOCKS
#define RG_MON_MIN_BLOCKS 16
#endif
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, RG_MON_MIN_BLOCKS)
k_step_monsters(DevBatch b, int auto_reset, int which) {
  unsigned char* const smem = rg_smem;
  const int parity = (int)(*b.dstep & 1u);
  TraceScope trace(b, which ? TK_MONSTERS_B : TK_MONSTERS);
  const uint32_t count = b.mon_count[which * 2 + parity];
  const uint32_t* const list = which ? b.mon_list_b : b.mon_list;
  const int warp = threadIdx.x >> 5;
  unsigned char* const base = smem + (size_t)warp * warp_smem(b);
  Stager sg = stager_init(b, base);
  // envs are handed out one at a time: their cost varies by an order of magnitude (a BFS or not),
  // and warps on SMs that also host a background generator block run slower
  for (;;) {
    uint32_t i = 0;
    if ((threadIdx.x & 31) == 0) i = atomicAdd(b.mon_count + 4 + which * 2 + parity, 1u);
    i = __shfl_sync(RG_FULL, i, 0);
    if (i >= count) break;
    const int64_t env = (int64_t)list[i];
    Ctx c;
    fill_ctx(b, c, sg, base, env, PL_BOTH);  // surface for the moves, both planes if the step ends with a compose
    EnvState* st = c.st;
    c.msg = st->f_msg;
    c.redraw = (st->f_flags & SF_REDRAW) ? 1u : 0u;
    c.status_upd = (st->f_flags & SF_STATUS) ? 1u : 0u;
    if (move_active_enemies(c)) st->ui_dead = 1;
    finish_env(b, c, env, auto_reset, parity);
    __syncwarp();
  }
}

// The whole step as one piece (with the floor generator), for the envs on the full-path list:
// descents, MoveUntil, and the reset half of a terminal step.
RG_DEV void step_env_full(const DevBatch& b, Ctx& c, int64_t env, const uint8_t* __restrict__ actions, int auto_reset,
                          bool reset_only, int parity) {
  EnvState* st = c.st;
  if (reset_only) {  // second half of a terminal step: finish_env left gold_before in reward[] and its error in error[]
    const uint32_t gold_before = (uint32_t)b.reward[env];
    uint8_t err = b.error[env];
    reset_env<false>(c);
    if (c.panic) {
      st->error = RG_ERR_PANIC;
      if (!err) err = RG_ERR_PANIC;
    }
    compose(c);
    const int32_t diff = (int32_t)st->status[1] - (int32_t)gold_before;
    emit_obs(b, c, env, diff > 0 ? diff : 0, err, 1);
    maybe_request_spec(b, c, env);
    close_env(b, c, env);
    return;
  }
  // k_step_scan has already checked the early-outs for this env; the step ends like every other one (finish_env:
  // a terminal step takes the prefetched game or goes to the reset pass, which runs after this kernel has joined)
  int d;
  const int act = map_key(actions[env], d);
  st->f_gold_before = st->status[1];
  process_action<false>(c, act, d);
  finish_env(b, c, env, auto_reset, parity);
}

// End of a step: advances the step counter and, on the steps that kick a background pass, fixes the
// window of refill requests that pass serves: [end of the previous window, current tail). Runs in one
// thread after every kernel of the step has finished.
RG_DEV void fix_window(const DevBatch& b) {  // window of background pass number dstep[3] (host: rg_batch::passes)
  const uint32_t p = b.dstep[3]++;
  uint32_t* win = b.refill_win + 2 * (p % 8u);
  win[0] = b.refill_ctl[2];
  win[1] = b.refill_ctl[2] = *reinterpret_cast<volatile uint32_t*>(b.refill_ctl);
}
RG_DEV void step_end(const DevBatch& b, int auto_reset) {
  if (b.spec) {  // the window of skeleton requests the pass kicked after this step serves
    uint32_t* win = b.spec_win + 2 * (b.dstep[0] % 8u);
    win[0] = b.spec_ctl[2];
    win[1] = b.spec_ctl[2] = *reinterpret_cast<volatile uint32_t*>(b.spec_ctl);
  }
  b.dstep[0] += 1u;
  if (!auto_reset || !b.prefetch) return;
  const uint32_t k = b.dstep[1]++;
  if (k % (uint32_t)b.prefetch_every) return;
  fix_window(b);
}
__global__ void k_step_end(DevBatch b, int auto_reset) {
  if (threadIdx.x == 0) step_end(b, auto_reset);
}
// Grid-stride over the full-path list; exits at once when the list is empty.
// With `finalize` the last block to finish also does the end-of-step bookkeeping (one kernel boundary less).
__global__ void __launch_bounds__(GEN_WPB * 32) k_step_gen(DevBatch b, const uint8_t* __restrict__ actions,
                                                                  int auto_reset, int resets, int finalize) {
  unsigned char* const smem = rg_smem;
  const int parity = (int)(*b.dstep & 1u);
  TraceScope trace(b, resets ? TK_RESETS : TK_FULL);
  const uint32_t count = resets ? b.reset_count[parity] : b.defer_count[parity];
  const uint32_t* const work = resets ? b.reset_list : b.defer_list;
  const int warp = threadIdx.x >> 5;
  unsigned char* const base = smem + (size_t)warp * warp_smem(b);
  Stager sg = stager_init(b, base);
  for (uint32_t i = blockIdx.x * GEN_WPB + warp; i < count; i += gridDim.x * GEN_WPB) {
    const uint32_t item = work[i];
    const int64_t env = (int64_t)(item & 0x7FFFFFFFu);
    Ctx c;
    fill_ctx(b, c, sg, base, env, (item & DEFER_RESET) ? PL_NONE : PL_BOTH);
    step_env_full(b, c, env, actions, auto_reset, (item & DEFER_RESET) != 0, parity);
    __syncwarp();
  }
  if (finalize) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(b.dstep + 2, 1u) == gridDim.x - 1) {  // every other block has finished (and read the parity)
        b.dstep[2] = 0;
        __threadfence();
        step_end(b, auto_reset);
      }
    }
  }
}

// Background generation of every env's NEXT episode (GameConfig::build for the seed its next reset
// will use - fixed, or already derived for `seed: null` - core/src/lib.rs:157-165,193-228). Runs on
// two alternating high-priority background streams concurrently with the step kernels, so the ~1000 dependent RNG
// draws of a floor are off the step's critical path; finish_env moves the finished game in
// when the episode ends. Ownership of a buffer is handed over through sp_state (0: this kernel
// may write it, 1: the step kernels may read it), with a fence on each side.
__global__ void __launch_bounds__(PF_MAX_WPB * 32, 1) k_prefetch(DevBatch b, int slot) {
  const uint32_t PF_WPB = blockDim.x >> 5;
  unsigned char* const smem = rg_smem;
  TraceScope trace(b, TK_PREFETCH);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned char* const base = smem + (size_t)warp * warp_smem(b);
  Stager sg = stager_init(b, base);
  const uint32_t begin = b.refill_win[2 * slot], end = b.refill_win[2 * slot + 1];
  for (uint32_t it = begin + blockIdx.x * PF_WPB + warp; (int32_t)(end - it) > 0; it += gridDim.x * PF_WPB) {
    const int64_t env = (int64_t)b.refill_ring[it & (b.refill_cap - 1u)];
    // passes overlap (one per step, two streams): one warp at a time per env; the loser hands the
    // request on to a later pass
    uint32_t got = 0;
    if (lane == 0) got = atomicCAS(b.sp_lock + env, 0u, 1u) == 0u ? 1u : 0u;
    got = __shfl_sync(RG_FULL, got, 0);
    if (!got) {
      if (lane == 0) {
        const uint32_t t = atomicAdd(b.refill_ctl, 1u);
        b.refill_ring[t & (b.refill_cap - 1u)] = (uint32_t)env;
      }
      continue;
    }
    __threadfence();
    const uint32_t e0 = *reinterpret_cast<volatile uint32_t*>(&b.st[env].episode);  // live episode counter
    // Episode e0+k is built from the state that precedes it: the live env (k = 1) or the
    // already-built game of episode e0+k-1 (its EnvState carries the seed the next reset uses).
    for (uint32_t k = 1; k <= (uint32_t)SP_DEPTH; ++k) {
      const int64_t sp = env * SP_DEPTH + (int64_t)((e0 + k) % SP_DEPTH);
      const bool have = *reinterpret_cast<volatile uint8_t*>(b.sp_state + sp) == 1 &&
                        *reinterpret_cast<volatile uint32_t*>(&b.sp_st[sp].episode) == e0 + k;
      if (have) continue;
      if (*reinterpret_cast<volatile uint8_t*>(b.sp_state + sp) == 1) break;  // holds a game the step kernels may still take
      if (*reinterpret_cast<volatile uint32_t*>(b.sp_cancel + sp) == e0 + k) break;  // that episode was built synchronously
      __threadfence();
      Ctx c;
      if (k == 1) {
        fill_ctx(b, c, sg, base, env, PL_NONE);
      } else {
        const int64_t prev = env * SP_DEPTH + (int64_t)((e0 + k - 1) % SP_DEPTH);
        DevBatch bb = b;
        bb.st = b.sp_st;  // basis = the previous prefetched game
        fill_ctx(bb, c, sg, base, prev, PL_NONE);
      }
      c.P = b.cfg_idx ? b.P + b.cfg_idx[env] : b.P;  // fill_ctx was given a ring slot, not the env, for k > 1
      if (c.st->episode != e0 + k - 1) break;  // the basis moved on under us: next pass
      c.g_screen = b.sp_screen + sp * b.CP;
      c.g_rows = nullptr;  // a prefetched game's screen is not the live one (the swap-in marks every row)
      c.g_hist = b.sp_hist + sp * b.HB;
      c.g_walk = b.sp_walk + sp * (int64_t)(b.H * b.WW);
      reset_env<true>(c);
      if (c.panic) c.st->error = RG_ERR_PANIC;
      compose(c);
      write_back(b, c, sp, true, true, b.sp_st, b.sp_surface, b.sp_attr);
      if (c.lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the writes themselves, not just the reads
      __threadfence();
      __syncwarp();
      // the step kernels may have given up on this game in the meantime (finish_env's miss path)
      const bool cancelled = *reinterpret_cast<volatile uint32_t*>(b.sp_cancel + sp) == e0 + k ||
                             *reinterpret_cast<volatile uint32_t*>(&b.st[env].episode) >= e0 + k;
      if (cancelled) break;
      if (c.lane == 0) *reinterpret_cast<volatile uint8_t*>(b.sp_state + sp) = 1;
      count_event(b, c, RGS_PREFETCH_BUILT);
      __syncwarp();
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(b.sp_lock + env, 0u);
  }
}

// Background builder of next-level skeletons (speculative descents). One warp per request: it snapshots the env's
// level and dungeon stream (a racy read of the live state - the consumer, take_spec, accepts a skeleton only if both
// still match when the descent happens), runs gen_skeleton for level + 1 in shared memory and publishes planes,
// rooms and the tag under the env's seqlock. A descent that finds no valid skeleton simply generates as before.
__global__ void __launch_bounds__(PF_MAX_WPB * 32, 1) k_spec_build(DevBatch b, int slot) {
  const uint32_t WPB = blockDim.x >> 5;
  unsigned char* const smem = rg_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per warp: the two planes and a state block like every warp kernel, then CP bytes for dig_maze's stack (passes
  // overlap on several streams, so the stack cannot live in a buffer indexed by the warp's number)
  unsigned char* const base = smem + (size_t)warp * (warp_smem(b) + (size_t)b.CP);
  const uint32_t gwarp = blockIdx.x * WPB + warp;
  const uint32_t begin = b.spec_win[2 * slot], end = b.spec_win[2 * slot + 1];
  for (uint32_t it = begin + gwarp; (int32_t)(end - it) > 0; it += gridDim.x * WPB) {
    const int64_t env = (int64_t)b.spec_ring[it & (b.spec_cap - 1u)];
    const int level = *reinterpret_cast<const volatile int32_t*>(&b.st[env].level);
    const uint4 r4 = __ldcg(reinterpret_cast<const uint4*>(b.st[env].rng));
    {  // already built for this very state (a repeated request)?
      const SpecTag* t = b.spec_tag + env;
      const uint4 have = __ldcg(reinterpret_cast<const uint4*>(t->rd_before));
      const int2 lv_ok = __ldcg(reinterpret_cast<const int2*>(&t->level));
      const uint32_t sq = *reinterpret_cast<const volatile uint32_t*>(b.spec_seq + env);
      if (!(sq & 1u) && lv_ok.y && lv_ok.x == level + 1 && have.x == r4.x && have.y == r4.y && have.z == r4.z && have.w == r4.w)
        continue;
    }
    uint32_t got = 0;  // one builder per env at a time (duplicate requests inside one window)
    if (lane == 0) got = atomicCAS(b.spec_lock + env, 0u, 1u) == 0u ? 1u : 0u;
    got = __shfl_sync(RG_FULL, got, 0);
    if (!got) continue;
    if (lane == 0) atomicAdd(b.spec_seq + env, 1u);  // odd: being written
    __threadfence();
    __syncwarp();
    Ctx c;
    c.soff = (uint32_t)(base - rg_smem);
    c.S = base;
    c.A = base + b.CP;
    c.st = reinterpret_cast<EnvState*>(base + 2 * (size_t)b.CP);
    c.col_room = b.room_lut;
    c.row_room = b.room_lut + 160;
    c.P = b.cfg_idx ? b.P + b.cfg_idx[env] : b.P;
    c.W = b.W; c.H = b.H; c.C = b.C; c.CP = b.CP; c.WW = b.WW;
    c.lane = lane;
    c.nx = b.nx; c.ny = b.ny;
    c.rsx = b.rsx; c.rsy = b.rsy;
    c.nrooms = c.nx * c.ny;
    c.g_screen = base + warp_smem(b);  // dig_maze's stack
    c.g_hist = nullptr; c.g_rows = nullptr; c.g_walk = nullptr; c.g_dist = nullptr; c.g_bfs = nullptr; c.g_wsnap = nullptr;
    c.g_spec_seq = nullptr; c.g_spec_hits = nullptr;
    c.redraw = c.status_upd = c.dead = c.msg = c.hist_done = c.a_dirty = c.s_dirty = c.panic = 0;
    if (lane < MAX_ROOMS) reinterpret_cast<uint2*>(c.st->rooms)[lane] = make_uint2(0u, 0u);
    Rng rd;
    rd.x = r4.x; rd.y = r4.y; rd.z = r4.z; rd.w = r4.w;
    gen_call::gen_skeleton(c, rd, (uint32_t)(level + 1));
    __syncwarp();
    if (lane < MAX_ROOMS) reinterpret_cast<uint2*>(b.spec_rooms + env * MAX_ROOMS)[lane] = reinterpret_cast<const uint2*>(c.st->rooms)[lane];
    if (lane == 0) {
      SpecTag* t = b.spec_tag + env;
      *reinterpret_cast<uint4*>(t->rd_before) = r4;
      *reinterpret_cast<uint4*>(t->rd_after) = make_uint4(rd.x, rd.y, rd.z, rd.w);
      t->level = level + 1;
      t->ok = c.panic ? 0u : 1u;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(b.spec_S + env * b.CP, smem_u32(c.S), (uint32_t)b.CP);
      bulk_s2g(b.spec_A + env * b.CP, smem_u32(c.A), (uint32_t)b.CP);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the writes themselves, not just the reads
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      atomicAdd(b.spec_seq + env, 1u);  // even: published
      __threadfence();
      atomicExch(b.spec_lock + env, 0u);
    }
    __syncwarp();
  }
}

// Dungeon::move_enemy with an always-false skip, for the known-answer test (rogue/mod.rs:566-578)
__global__ void k_test_move_enemy(DevBatch b, int64_t env_id, int fx, int fy, int tx, int ty, int* out3) {
  unsigned char* const smem = rg_smem;
  Ctx c;
  const int64_t env = env_id;
  if ((threadIdx.x >> 5) != 0) return;
  Stager sg = stager_init(b, smem);
  fill_ctx(b, c, sg, smem, env, PL_BOTH);
  int ox = -1, oy = -1;
  int kind = move_enemy(c, fx, fy, tx, ty, 0, true, ox, oy);
  if (c.lane == 0) {
    out3[0] = c.panic ? -1 : kind;
    out3[1] = ox;
    out3[2] = oy;
  }
  close_env(b, c, env);
}

// Parity harness: finish every suspended DistCache map so that rg_dump_env can show whole maps.
// Does not change anything observable (a finished map holds the values the reference holds).
__global__ void __launch_bounds__(32) k_complete_maps(DevBatch b, int64_t env_lo, int64_t env_hi) {
  unsigned char* const smem = rg_smem;
  const int64_t env = env_lo + blockIdx.x;
  if (env >= env_hi) return;
  Ctx c;
  Stager sg = stager_init(b, smem);
  fill_ctx(b, c, sg, smem, env, PL_NONE);
  complete_all_maps(c);
  store_state(b, c, env);  // only the cache directory changed
}

// ---------------------------------------------------------------- observation encoders
// PlayerState::{gray,symbol}_image[_with_hist] python/src/lib.rs:72-111,158-205 and
// symbol::construct_symbol_map core/src/symbol.rs:51-71. One block per (env, plane group);
// every thread produces float4s, so the store stream is fully coalesced. HBM-write bound.
RG_DEV int tile_sym(uint32_t t) {  // Symbol::from_tile symbol.rs:17-40
  switch (t) {
    case ' ': return 0;
    case '@': return 1;
    case '#': return 2;
    case '.': return 3;
    case '-':
    case '|': return 4;
    case '%': return 5;
    case '+': return 6;
    case '^': return 7;
    case '!': return 8;
    case '?': return 9;
    case ']': return 10;
    case ')': return 11;
    case '/': return 12;
    case '*': return 13;
    case ':': return 14;
    case '=': return 15;
    case ',': return 16;
    default: return (t >= 'A' && t <= 'Z') ? (int)t - 'A' + 17 : -1;
  }
}

__global__ void __launch_bounds__(256) k_encode(DevBatch b, int mode, uint32_t flag, int with_hist, int channels,
                                               float* __restrict__ out) {
  const int64_t env = blockIdx.x;
  const int C = b.C;  // C is a multiple of 4 whenever W is; the tail is handled per element
  const uint8_t* scr = b.screen + env * b.CP;
  const int symbols = (int)b.P->symbols;
  const int base = mode == 0 ? 1 : symbols;
  float* o = out + env * (int64_t)channels * C;
  __shared__ uint8_t sym[160 * 48];
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    int s = tile_sym(scr[i]);
    if (s < 0 || (mode == 1 && s >= symbols - 1)) {  // InvalidTileError symbol.rs:60-64
      bad = 1;
      s = 0;
    }
    sym[i] = (uint8_t)s;
  }
  __syncthreads();
  if (C & 3) {  // odd sizes: plain per-element path
    const int order_s[9] = {0, 2, 3, 4, 5, 6, 7, 8, 9};
    const uint32_t* stv = b.status + env * 10;
    for (int64_t i = threadIdx.x; i < (int64_t)channels * C; i += blockDim.x) {
      int ch = (int)(i / C), cell = (int)(i - (int64_t)ch * C);
      float v;
      if (ch < base) {
        v = mode == 0 ? (float)sym[cell] / (float)symbols : ((ch < symbols - 1 && sym[cell] == ch) ? 1.f : 0.f);
      } else {
        int k = ch - base, bit = 0, seen = 0;
        for (bit = 0; bit < 9; ++bit)
          if (flag & (1u << bit)) {
            if (seen == k) break;
            ++seen;
          }
        if (bit < 9) v = (float)(int32_t)stv[order_s[bit]];
        else v = ((b.hist + env * b.HB)[cell >> 3] >> (cell & 7)) & 1u ? 1.f : 0.f;
      }
      o[i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0 && bad) {
      b.error[env] = RG_ERR_SETTING;
      atomicOr(b.errflag, 1u << RG_ERR_SETTING);
    }
    return;
  }
  const int C4 = C / 4;
  if (mode == 0) {
    const float inv = (float)symbols;
    for (int i = threadIdx.x; i < C4; i += blockDim.x) {
      float4 v = make_float4((float)sym[4 * i] / inv, (float)sym[4 * i + 1] / inv, (float)sym[4 * i + 2] / inv,
                             (float)sym[4 * i + 3] / inv);
      reinterpret_cast<float4*>(o)[i] = v;
    }
  } else {
    // channels 0 .. symbols-2 are one-hot; channel symbols-1 stays zero
    const int total = symbols * C4;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      int ch = i / C4, q = i - ch * C4;
      float4 v = make_float4(sym[4 * q] == ch ? 1.f : 0.f, sym[4 * q + 1] == ch ? 1.f : 0.f,
                             sym[4 * q + 2] == ch ? 1.f : 0.f, sym[4 * q + 3] == ch ? 1.f : 0.f);
      if (ch == symbols - 1) v = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(o)[i] = v;
    }
  }
  // StatusFlagInner::copy_status python/src/flags.rs:87-115
  int off = base;
  const uint32_t* st = b.status + env * 10;
  const int order[9] = {0, 2, 3, 4, 5, 6, 7, 8, 9};
  for (int bit = 0; bit < 9; ++bit)
    if (flag & (1u << bit)) {
      const float v = (float)(int32_t)st[order[bit]];
      float4* p = reinterpret_cast<float4*>(o + (int64_t)off * C);
      for (int i = threadIdx.x; i < C4; i += blockDim.x) p[i] = make_float4(v, v, v, v);
      ++off;
    }
  if (with_hist) {  // copy_hist python/src/lib.rs:105-111
    const uint8_t* hb = b.hist + env * b.HB;
    float4* p = reinterpret_cast<float4*>(o + (int64_t)off * C);
    for (int i = threadIdx.x; i < C4; i += blockDim.x) {
      uint32_t byte = hb[(4 * i) >> 3];
      int sh = (4 * i) & 7;
      p[i] = make_float4((byte >> sh) & 1u ? 1.f : 0.f, (byte >> (sh + 1)) & 1u ? 1.f : 0.f,
                         (byte >> (sh + 2)) & 1u ? 1.f : 0.f, (byte >> (sh + 3)) & 1u ? 1.f : 0.f);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && bad) {
    b.error[env] = RG_ERR_SETTING;
    atomicOr(b.errflag, 1u << RG_ERR_SETTING);
  }
}

// Gray images (mode 0) have 1 + S (+1) planes only: one block per env leaves the GPU launch-bound
// (7.7 KB of output per block). Flat version: a thread owns one float4 position (plane, 4 cells) and
// writes it for GRAY_ENVS consecutive envs, so the index arithmetic is done once and consecutive threads
// write consecutive float4s of the same plane. Same values as k_encode (python/src/lib.rs:72-111).
constexpr int GRAY_ENVS = 8;
__global__ void __launch_bounds__(256) k_encode_gray(DevBatch b, uint32_t flag, int channels, float* __restrict__ out) {
  const int C = b.C, C4 = C / 4;  // only launched when C % 4 == 0
  const int per_env = channels * C4;
  // tile byte -> gray level (Symbol::from_tile / symbols, the same f32 division as k_encode), -1 = no symbol
  __shared__ float gray_of_tile[256];
  {
    const int sym = tile_sym(threadIdx.x);
    gray_of_tile[threadIdx.x] = sym < 0 ? -1.f : (float)sym / (float)b.P->symbols;
  }
  __syncthreads();
  const int r = (int)(blockIdx.y * blockDim.x + threadIdx.x);
  if (r >= per_env) return;
  const int ch = r / C4, q = r - ch * C4;
  const int nstat = __popc(flag & 0x1FFu);
  int sidx = 0;
  if (ch >= 1 && ch <= nstat) {  // StatusFlagInner::copy_status python/src/flags.rs:87-115: the (ch-1)-th set bit
    const int bit = nth_set_bit(flag & 0x1FFu, (uint32_t)(ch - 1));
    sidx = bit == 0 ? 0 : bit + 1;
  }
  const int64_t env0 = (int64_t)blockIdx.x * GRAY_ENVS;
#pragma unroll 4
  for (int e = 0; e < GRAY_ENVS; ++e) {
    const int64_t env = env0 + e;
    if (env >= b.n) break;
    float4 v;
    if (ch == 0) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(b.screen + env * b.CP + 4 * q);
      v = make_float4(gray_of_tile[w & 0xFFu], gray_of_tile[(w >> 8) & 0xFFu], gray_of_tile[(w >> 16) & 0xFFu],
                      gray_of_tile[w >> 24]);
      if (v.x < 0.f || v.y < 0.f || v.z < 0.f || v.w < 0.f) {  // InvalidTileError symbol.rs:60-64
        b.error[env] = RG_ERR_SETTING;
        atomicOr(b.errflag, 1u << RG_ERR_SETTING);
        v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
      }
    } else if (ch <= nstat) {
      const float f = (float)(int32_t)b.status[env * 10 + sidx];
      v = make_float4(f, f, f, f);
    } else {  // copy_hist python/src/lib.rs:105-111
      const uint32_t byte = (b.hist + env * b.HB)[(4 * q) >> 3];
      const int sh = (4 * q) & 7;
      v = make_float4((byte >> sh) & 1u ? 1.f : 0.f, (byte >> (sh + 1)) & 1u ? 1.f : 0.f, (byte >> (sh + 2)) & 1u ? 1.f : 0.f,
                      (byte >> (sh + 3)) & 1u ? 1.f : 0.f);
    }
    reinterpret_cast<float4*>(out)[env * per_env + r] = v;
  }
}

// Compact observation (rg_encode_compact): symbol ids, not one-hot planes. A thread maps 16 cells (one 128-bit load,
// one 128-bit store) through a 256-entry table in shared memory; HBM-bound at 2 bytes per cell.
__global__ void __launch_bounds__(256) k_encode_compact(DevBatch b, uint8_t* __restrict__ sym_out, int32_t* __restrict__ status_out,
                                                       uint8_t* __restrict__ hist_out) {
  __shared__ uint8_t sym_of_tile[256];
  {
    const int sym = tile_sym(threadIdx.x);
    sym_of_tile[threadIdx.x] = sym < 0 ? 0xFFu : (uint8_t)sym;  // no symbol at all: InvalidTileError symbol.rs:60-64
  }
  __syncthreads();
  const int pieces = b.CP / 16;
  const int64_t total = b.n * pieces;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t env = i / pieces;
    const int piece = (int)(i - env * pieces), base = piece * 16;
    const uint4 v = *reinterpret_cast<const uint4*>(b.screen + env * b.CP + base);
    const uint32_t in[4] = {v.x, v.y, v.z, v.w};
    uint32_t out[4];
    bool bad = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t o = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t sy = sym_of_tile[(in[q] >> (8 * k)) & 0xFFu];
        if (sy == 0xFFu) {
          if (base + q * 4 + k < b.C) bad = true;
          sy = 0;
        }
        o |= sy << (8 * k);
      }
      out[q] = o;
    }
    if (bad) {
      b.error[env] = RG_ERR_SETTING;
      atomicOr(b.errflag, 1u << RG_ERR_SETTING);
    }
    uint8_t* dst = sym_out + env * (int64_t)b.C + base;
    if ((b.C & 15) == 0) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {  // odd sizes: the dense rows are not 16-byte aligned
      for (int k = 0; k < 16 && base + k < b.C; ++k) dst[k] = (uint8_t)(out[k >> 2] >> (8 * (k & 3)));
    }
    if (hist_out) {  // Floor::history_map as 0/1 bytes
      const uint32_t bits = reinterpret_cast<const uint16_t*>(b.hist + env * b.HB)[piece];
      uint8_t* hd = hist_out + env * (int64_t)b.C + base;
      if ((b.C & 15) == 0) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t nib = (bits >> (4 * q)) & 0xFu;
          w[q] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        }
        *reinterpret_cast<uint4*>(hd) = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        for (int k = 0; k < 16 && base + k < b.C; ++k) hd[k] = (uint8_t)((bits >> k) & 1u);
      }
    }
    if (status_out && piece == 0) {  // StatusFlagInner::to_vector order python/src/flags.rs:63-85
      const uint32_t* st = b.status + env * 10;
      int32_t* so = status_out + env * 9;
      so[0] = (int32_t)st[0];
#pragma unroll
      for (int k = 1; k < 9; ++k) so[k] = (int32_t)st[k + 1];
    }
  }
}

// ---------------------------------------------------------------- trainer-facing step (rg_step_train)
// Before the step: gym action index -> ASCII key (RogueEnv.ACTIONS, python/rogue_gym/envs/rogue_env.py:159-172).
__global__ void __launch_bounds__(256) k_keys_from_index(DevBatch b, const void* __restrict__ idx, int index_bytes,
                                                         uint8_t* __restrict__ keys) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= b.n) return;
  if (index_bytes == 0) {  // already ASCII keys
    keys[env] = reinterpret_cast<const uint8_t*>(idx)[env];
    return;
  }
  long long a = index_bytes == 8 ? reinterpret_cast<const long long*>(idx)[env]
              : index_bytes == 4 ? (long long)reinterpret_cast<const int*>(idx)[env]
                                 : (long long)reinterpret_cast<const uint8_t*>(idx)[env];
  // an index outside the table becomes a key outside the key map: the env reports InvalidInput, as
  // ParallelRogueEnv.step raises for it
  keys[env] = (a >= 0 && a < 11) ? (uint8_t)".hjklnbuy>s"[a] : (uint8_t)'?';
}
// After the step: float reward = gold gained (+ stair bonus when the env is deeper than the level remembered
// for it; the remembered level follows the env, so it is 1 again after an auto-reset:
// StairRewardParallel, python/rogue_gym/envs/wrappers.py:44-64).
__global__ void __launch_bounds__(256) k_train_reward(DevBatch b, float stair_reward, int32_t* __restrict__ level_seen,
                                                      float* __restrict__ reward_out) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= b.n) return;
  float r = (float)b.reward[env];
  const int32_t level = (int32_t)b.status[env * 10];
  if (level > level_seen[env]) r += stair_reward;
  level_seen[env] = level;
  reward_out[env] = r;
}

// per-env state hash for large-N parity checks (matches oracle orc_state_hash)
__global__ void k_state_hash(DevBatch b, uint64_t* out) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= b.n) return;
  uint64_t h = 0xcbf29ce484222325ull;
  auto mix = [&](uint8_t v) { h ^= v; h *= 0x100000001b3ull; };
  const uint8_t* scr = b.screen + env * b.CP;
  for (int i = 0; i < b.C; ++i) mix(scr[i]);
  const EnvState* st = b.st + env;
  for (int i = 0; i < 10; ++i)
    for (int k = 0; k < 4; ++k) mix((uint8_t)(st->status[i] >> (8 * k)));
  for (int i = 0; i < 12; ++i)
    for (int k = 0; k < 4; ++k) mix((uint8_t)(st->rng[i] >> (8 * k)));
  int32_t s[6] = {st->px, st->py, st->hp, st->level, (int32_t)st->steps, (int32_t)st->is_terminal};
  for (int i = 0; i < 6; ++i)
    for (int k = 0; k < 4; ++k) mix((uint8_t)((uint32_t)s[i] >> (8 * k)));
  out[env] = h;
}

// parity harness (rg_export_floors): room table + player position of every env
__global__ void k_export_rooms(DevBatch b, int16_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n * MAX_ROOMS) return;
  const int64_t env = i / MAX_ROOMS;
  const int r = (int)(i % MAX_ROOMS);
  const EnvState* st = b.st + env;
  int16_t* o = out + i * 8;
  const bool live = r < b.nx * b.ny;
  const RoomD rm = st->rooms[live ? r : 0];
  o[0] = live ? rm.kind : -1;
  o[1] = live && (rm.flags & RF_DARK) ? 1 : 0;
  o[2] = live ? rm.x0 : 0;
  o[3] = live ? rm.y0 : 0;
  o[4] = live ? rm.x1 : 0;
  o[5] = live ? rm.y1 : 0;
  o[6] = st->px;
  o[7] = st->py;
}

// the state's own terminal flag (Instruction::State, thread_impls.rs:131), as opposed to done[] = what the last step returned
__global__ void k_state_terminal(DevBatch b, uint8_t* __restrict__ out) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env < b.n) out[env] = b.st[env].is_terminal;
}

// Instruction::Seed for every env (python/src/thread_impls.rs:125-128): stored, used by the next reset
__global__ void k_seed(DevBatch b, const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi, int seeded, int64_t count) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= count) return;
  EnvState* st = b.st + env;
  const uint64_t l = lo[env], h = hi ? hi[env] : 0ull;
  st->seed[0] = (uint32_t)l;
  st->seed[1] = (uint32_t)(l >> 32);
  st->seed[2] = (uint32_t)h;
  st->seed[3] = (uint32_t)(h >> 32);
  st->seeded = (uint8_t)seeded;
}

// PlayerState.history as 0/1 bytes [N][C] for host consumers (python/src/lib.rs:33)
__global__ void k_unpack_hist(DevBatch b, uint8_t* __restrict__ out) {
  const int64_t env = blockIdx.x;
  const uint8_t* hb = b.hist + env * b.HB;
  uint8_t* o = out + env * (int64_t)b.C;
  for (int i = threadIdx.x; i < b.C; i += blockDim.x) o[i] = (hb[i >> 3] >> (i & 7)) & 1u;
}

// ---------------------------------------------------------------- host mirror (delta write-back)
// The host-facing step (what the PyO3 layer does per call: Vec<PlayerState> out, python/src/lib.rs:
// 315-321) has to leave the whole observation block in host memory. Copying it costs 129 MB per step
// at 65 536 envs (PCIe-bound, ~2.4 ms) although a step changes a few cells per screen. The mirror is
// pinned host memory mapped into the device address space; these kernels compare the device block with
// a device-resident shadow of what the host already holds and store only the 16-byte pieces that
// differ - to the shadow and, over PCIe, straight into the host buffer (measured: ~70 us of a step's
// ~135 us mirror cost is the ~55 k small PCIe writes themselves).

RG_DEV bool differs(const uint4& a, const uint4& b) { return ((a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z) | (a.w ^ b.w)) != 0u; }

// One warp per env (grid-stride). Lanes 0-12 first check the env's scalars (status, reward, message,
// done | error) against the shadow; then, compose() having recorded which rows of the screen / visited
// map it rewrote (scr_rows), only the 16-byte pieces that overlap those rows are compared at all.
// `pass`: 0 every env; 1 only the envs whose step was finished by the player kernel (full_path[] == 0) -
// this pass runs beside the monster, full-path and reset kernels, which own the other envs; 2 only those others.
__global__ void __launch_bounds__(256) k_mirror(DevBatch b, MirrorArgs m, int pass, int64_t env_lo, int64_t env_hi) {
  // (the second pass runs after the step counter has advanced: it reports into the slot of the step it belongs to)
  DevBatch tb = b;
  uint32_t prev_step = *b.dstep - (pass == 2 ? 1u : 0u);
  tb.dstep = &prev_step;
  TraceScope trace(tb, pass == 2 ? TK_MIRROR2 : TK_MIRROR1);
  const int lane = threadIdx.x & 31;
  const int n_scr = b.CP / 16, n_hist = m.with_hist ? b.HB / 16 : 0;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t sent = 0;
  for (int64_t env = env_lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); env < env_hi; env += warps) {
    if (pass) {
      const bool early = b.full_path[env] == 0;
      if (early != (pass == 1)) continue;
    }
    const uint64_t rows = b.scr_rows[env];
    if (!rows) continue;
    // `wide`: a changed 16-byte piece of the screen takes its whole 64-byte line along (the four lanes of a group
    // store together). The host then receives full-line writes: 8 GPUs writing 16-byte pieces into one host's memory
    // share a budget of ~1.1 G partial-line writes per second (measured), which made the host-facing step 1.6x slower
    // at 8 GPUs than at 1.
    const bool wide = m.wide && (b.C & 63) == 0;
    for (int base = 0; base < n_scr + n_hist; base += 32) {
      const int piece = base + lane;
      const bool is_screen = piece < n_scr;
      const int off = (is_screen ? piece : piece - n_scr) * 16;
      // cells covered: 16 per screen piece, 128 per piece of the bit-packed visited map
      const int c0 = is_screen ? off : off * 8, c1 = min((is_screen ? off + 15 : off * 8 + 127), b.C - 1);
      bool active = piece < n_scr + n_hist && c0 < b.C;
      if (active) {
        const int r0 = c0 / b.W, r1 = c1 / b.W;
        active = ((rows >> r0) & ((2ull << (r1 - r0)) - 1ull)) != 0ull;
      }
      const uint32_t group = 0xFu << (lane & ~3);
      const bool grouped = wide && is_screen;
      const uint32_t active_m = __ballot_sync(RG_FULL, active);  // (every lane votes: no short-circuit around a ballot)
      const bool look = active || (grouped && (active_m & group));
      const uint8_t* cur_p = is_screen ? b.screen + env * b.CP + off : b.hist + env * b.HB + off;
      uint8_t* sh_p = is_screen ? m.s_screen + env * b.CP + off : m.s_hist + env * b.HB + off;
      uint4 cur = make_uint4(0, 0, 0, 0);
      bool changed = false;
      if (look) {
        cur = *reinterpret_cast<const uint4*>(cur_p);
        changed = differs(cur, *reinterpret_cast<const uint4*>(sh_p));
      }
      const uint32_t changed_m = __ballot_sync(RG_FULL, changed);
      if (changed) *reinterpret_cast<uint4*>(sh_p) = cur;
      if (!(changed || (grouped && look && (changed_m & group)))) continue;
      if (!is_screen) {
        *reinterpret_cast<uint4*>(m.h_hist + env * b.HB + off) = cur;
        sent += 16;
      } else if ((b.C & 15) == 0) {
        *reinterpret_cast<uint4*>(m.h_screen + env * (int64_t)b.C + off) = cur;
        sent += 16;
      } else {  // odd screen sizes: the dense host rows are not 16-byte aligned
        const uint8_t* cb = reinterpret_cast<const uint8_t*>(&cur);
        for (int k = 0; k < 16 && off + k < b.C; ++k) m.h_screen[env * (int64_t)b.C + off + k] = cb[k];
        sent += (uint32_t)min(16, b.C - off);
      }
    }
    __syncwarp();
    if (lane == 0) b.scr_rows[env] = 0ull;
  }
  if (pass != 1) {
    // The small arrays of the observation block (status, reward, message, done, error), for every env at once, by the
    // pass that ends the call: flat 16-byte pieces against flat shadows, a changed piece sent with its 64-byte line.
    // (16 envs share a line of `message`: one full-line write replaces up to 16 four-byte partial writes.)
    const uint8_t* src[5] = {reinterpret_cast<const uint8_t*>(b.status), reinterpret_cast<const uint8_t*>(b.reward),
                             reinterpret_cast<const uint8_t*>(b.message), b.done, b.error};
    uint8_t* dst[5] = {reinterpret_cast<uint8_t*>(m.h_status), reinterpret_cast<uint8_t*>(m.h_reward),
                       reinterpret_cast<uint8_t*>(m.h_message), m.h_done, m.h_error};
    const int64_t bytes[5] = {b.n * 40, b.n * 4, b.n * 4, b.n, b.n};
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t sh_off = 0;
    const uint32_t group = 0xFu << (lane & ~3);
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      const int64_t pieces = (bytes[a] + 63) / 64 * 4;  // whole lines (the arrays and their shadows are padded)
      for (int64_t p0 = gwarp * 32; p0 < pieces; p0 += warps * 32) {
        const int64_t p = p0 + lane;
        bool changed = false;
        uint4 cur = make_uint4(0, 0, 0, 0);
        if (p < pieces) {
          cur = *reinterpret_cast<const uint4*>(src[a] + p * 16);
          if (p * 16 + 16 > bytes[a]) {  // the tail beyond the array: never compare or send garbage
            uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
            for (int k = 0; k < 16; ++k)
              if (p * 16 + k >= bytes[a]) w[k >> 2] &= ~(0xFFu << (8 * (k & 3)));
            cur = make_uint4(w[0], w[1], w[2], w[3]);
          }
          changed = differs(cur, *reinterpret_cast<const uint4*>(m.s_flat + sh_off + p * 16));
        }
        const uint32_t changed_m = __ballot_sync(RG_FULL, changed);
        if (changed) *reinterpret_cast<uint4*>(m.s_flat + sh_off + p * 16) = cur;
        if (p < pieces && (changed_m & group)) {
          *reinterpret_cast<uint4*>(dst[a] + p * 16) = cur;
          sent += 16;
        }
      }
      sh_off += pieces * 16;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sent += __shfl_xor_sync(RG_FULL, sent, o);
  if (lane == 0 && sent) atomicAdd(m.bytes, (unsigned long long)sent);
  if (pass != 1) {  // the pass that ends the call: the last block to finish hands the counters to the host
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(m.ticket, 1u) == gridDim.x - 1) {
        __threadfence();
        *m.h_bytes = *reinterpret_cast<volatile unsigned long long*>(m.bytes);
        *m.h_errflag = *reinterpret_cast<volatile uint32_t*>(b.errflag);
        *b.errflag = 0u;  // reported: rg_sync clears it the same way
        *m.bytes = 0ull;
        *m.ticket = 0u;
        __threadfence_system();
        const unsigned long long seq = *m.seq + 1ull;
        *m.seq = seq;
        *reinterpret_cast<volatile unsigned long long*>(m.h_seq) = seq;  // everything above is in host memory before this
        __threadfence_system();
      }
    }
  }
}

// ---------------------------------------------------------------- host mirror by lines (confined to a few SMs)
// Stores into mapped host memory back up behind PCIe, and an SM's memory pipeline serves its requests in order: while
// k_mirror's stores drain, every other kernel on the same SM waits behind them (timeline: the monster kernel beside the
// pass 45 -> 100 us, the player kernel 13 -> 40 us). With that pass confined to 16 SMs the other kernels run at full
// speed - but its one-warp-per-env loops of dependent loads then take 540 us. k_mirror_lines does the same compare with
// every thread busy (a work list of candidate 64-byte lines per chunk of envs), so that a few SMs (RG_MIRROR_SMS, one
// 1024-thread block each) are enough to keep PCIe full, and only those SMs see the back-pressure. The pass is bound by
// the number of PCIe writes (~45 k lines in 65-70 us, whether a write carries 64, 32 or 16 bytes). Tried and dropped
// (profiles/r2_mirror_modes.log): a line log in device memory written by every SM and streamed out by a few blocks
// (the extra kernel boundary costs more than the compare on few SMs), k_step_fast in two launches with the first
// half's pass beside the second (the split costs what the earlier start gains), 32-byte half-line writes.
// Used when the screen is a whole number of 64-byte lines per env; other sizes keep k_mirror.
enum : uint32_t { LINE_PIECE = 0x80000000u };  // line index flag: one 16-byte piece (visited map), index in 16-byte units

// A line has to leave as ONE 64-byte write (four adjacent lanes, 16 bytes each, in the same store instruction): a
// thread storing its own line piece by piece would send four partial-line writes (measured: the pass 2x slower). So
// the warp transposes: eight changed lines per round, lane l stores piece l & 3 of the (l >> 2)-th of them.
// Called by whole warps.
RG_DEV void emit_line(const MirrorArgs& m, bool changed, uint32_t off, const uint4 (&data)[4], uint32_t& sent) {
  const bool piece = (off & LINE_PIECE) != 0u;
  const int lane = threadIdx.x & 31;
  if (changed && piece) {
    *reinterpret_cast<uint4*>(m.h_base + (size_t)(off & ~LINE_PIECE) * 16) = data[0];
    sent += 16;
  }
  uint32_t todo = __ballot_sync(RG_FULL, changed && !piece);
  while (todo) {
    const uint32_t src = __fns(todo, 0, (lane >> 2) + 1);  // the lane that holds "my" line (0xFFFFFFFF: fewer than that are left)
    const int from = src == 0xFFFFFFFFu ? 0 : (int)src;
    const int q = lane & 3;
    uint4 v = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t x = __shfl_sync(RG_FULL, data[k].x, from), y = __shfl_sync(RG_FULL, data[k].y, from);
      const uint32_t z = __shfl_sync(RG_FULL, data[k].z, from), w = __shfl_sync(RG_FULL, data[k].w, from);
      if (k == q) v = make_uint4(x, y, z, w);
    }
    const uint32_t line = __shfl_sync(RG_FULL, off, from);
    if (src != 0xFFFFFFFFu) {
      reinterpret_cast<uint4*>(m.h_base + (size_t)line * 64)[q] = v;
      sent += 16;
    }
    const uint32_t eighth = __fns(todo, 0, 8);  // drop the eight lowest set bits
    todo = eighth == 0xFFFFFFFFu ? 0u : (eighth == 31u ? 0u : todo & (0xFFFFFFFFu << (eighth + 1)));
  }
}

// Chunks of `epb` envs, grid-stride. Phase 1, one thread per env: which 64-byte lines of the screen (16-byte pieces of
// the visited map) overlap the rows the step rewrote -> a work list in shared memory. Phase 2, one thread per list
// entry: compare with the shadow, update it, send. Then (`flat`) the small arrays (status, reward, message, done,
// error), one thread per line. `pass`: 1 = beside the step's other kernels: the envs k_step_fast finished, taken from
// positions [pos_lo, pos_hi) of its list of moves (the only envs that kernel redraws); 2 = after the step, 0 = on
// request (rg_mirror_sync): every env that still has rows marked. `publish`: the last block hands byte counter and
// error flag to the host.
__global__ void __launch_bounds__(1024) k_mirror_lines(DevBatch b, MirrorArgs m, int pass, int epb, int env_chunks, int flat,
                                                       int publish, int64_t pos_lo, int64_t pos_hi) {
  DevBatch tb = b;
  uint32_t prev_step = *b.dstep - (pass == 2 ? 1u : 0u);
  tb.dstep = &prev_step;
  TraceScope trace(tb, pass == 2 ? TK_MIRROR2 : TK_MIRROR1);
  extern __shared__ uint32_t smem_u32[];
  uint32_t* const envs = smem_u32;         // [epb] the chunk's envs
  uint32_t* const items = smem_u32 + epb;  // (index into envs << 16) | line, or | 0x8000 | piece of the visited map
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t total_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpe = b.C >> 6, ppe = m.with_hist ? b.HB >> 4 : 0;
  const uint32_t scr_line0 = (uint32_t)((m.h_screen - m.h_base) >> 6), hist_piece0 = (uint32_t)((m.h_hist - m.h_base) >> 4);
  uint32_t sent = 0;
  uint4 cur[4];
  if (pass == 1) pos_hi = min(pos_hi, (int64_t)b.fast_count[4]);  // the moves come first in the list
  for (int chunk = blockIdx.x; chunk < env_chunks; chunk += gridDim.x) {
    int64_t env = (int64_t)chunk * epb + threadIdx.x;
    if (pass == 1) {
      const int64_t pos = pos_lo + env;
      if (pos_lo + (int64_t)chunk * epb >= pos_hi) break;  // (block-uniform)
      env = (int)threadIdx.x < epb && pos < pos_hi ? (int64_t)b.fast_list_m[pos] : b.n;
    }
    if ((int)threadIdx.x < epb) envs[threadIdx.x] = (uint32_t)min(env, b.n);
    uint64_t rows = 0;
    if ((int)threadIdx.x < epb && env < b.n) {
      // pass 1: what k_step_fast finished (its leftovers are with the player kernel right now); passes 2 and 0 run after
      // everything else and take every env that still has rows marked (also from steps that were not mirrored)
      if (pass != 1 || b.full_path[env] == 0) {
        rows = b.scr_rows[env];
        if (rows) b.scr_rows[env] = 0ull;
        if (b.H < 64) rows &= (1ull << b.H) - 1ull;
      }
    }
    // lines / pieces are taken in ascending row order; consecutive rows share their boundary line: `last` dedupes
    auto walk = [&](uint32_t* out) -> uint32_t {
      uint32_t n = 0;
      int last_l = -1, last_p = -1;
      for (uint64_t r = rows; r; r &= r - 1) {
        const int y = __ffsll((long long)r) - 1;
        const int c0 = y * b.W, c1 = c0 + b.W - 1;
        for (int j = max(c0 >> 6, last_l + 1); j <= min(c1 >> 6, lpe - 1); ++j) {
          if (out) out[n] = (threadIdx.x << 16) | (uint32_t)j;
          ++n;
          last_l = j;
        }
        if (ppe)
          for (int p = max(c0 >> 7, last_p + 1); p <= min(c1 >> 7, ppe - 1); ++p) {
            if (out) out[n] = (threadIdx.x << 16) | 0x8000u | (uint32_t)p;
            ++n;
            last_p = p;
          }
      }
      return n;
    };
    const uint32_t cnt = walk(nullptr);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(RG_FULL, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (threadIdx.x == blockDim.x - 1) total_s = before + incl;
    if (cnt) walk(items + before + incl - cnt);
    __syncthreads();
    const uint32_t total = total_s;
    for (uint32_t i0 = 0; i0 < total; i0 += blockDim.x) {
      const uint32_t i = i0 + threadIdx.x;
      bool changed = false;
      uint32_t off = 0;
      if (i < total) {
        const uint32_t it = items[i];
        const int64_t e = (int64_t)envs[it >> 16];
        if (it & 0x8000u) {
          const uint32_t p = it & 0x7FFFu;
          cur[0] = *reinterpret_cast<const uint4*>(b.hist + e * b.HB + p * 16);
          uint4* sh = reinterpret_cast<uint4*>(m.s_hist + e * b.HB + p * 16);
          changed = differs(cur[0], *sh);
          if (changed) *sh = cur[0];
          off = LINE_PIECE | (hist_piece0 + (uint32_t)(e * (b.HB >> 4)) + p);
        } else {
          const uint32_t j = it & 0x7FFFu;
          const uint4* cp = reinterpret_cast<const uint4*>(b.screen + e * b.CP + j * 64);
          uint4* sh = reinterpret_cast<uint4*>(m.s_screen + e * b.CP + j * 64);
#pragma unroll
          for (int k = 0; k < 4; ++k) cur[k] = cp[k];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (differs(cur[k], sh[k])) { sh[k] = cur[k]; changed = true; }
          off = scr_line0 + (uint32_t)(e * lpe) + j;
        }
      }
      emit_line(m, changed, off, cur, sent);
    }
    __syncthreads();  // the work list is reused by the next chunk
  }
  if (flat) {
    const uint8_t* src[5] = {reinterpret_cast<const uint8_t*>(b.status), reinterpret_cast<const uint8_t*>(b.reward),
                             reinterpret_cast<const uint8_t*>(b.message), b.done, b.error};
    uint8_t* dst[5] = {reinterpret_cast<uint8_t*>(m.h_status), reinterpret_cast<uint8_t*>(m.h_reward),
                       reinterpret_cast<uint8_t*>(m.h_message), m.h_done, m.h_error};
    const int64_t bytes[5] = {b.n * 40, b.n * 4, b.n * 4, b.n, b.n};
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    int64_t sh_off = 0;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      const int64_t lines = (bytes[a] + 63) / 64;  // whole lines (the arrays and their shadows are padded)
      for (int64_t l0 = 0; l0 < lines; l0 += stride) {
        const int64_t l = l0 + t0;
        bool changed = false;
        if (l < lines) {
          const uint4* cp = reinterpret_cast<const uint4*>(src[a] + l * 64);
          uint4* sh = reinterpret_cast<uint4*>(m.s_flat + sh_off + l * 64);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            cur[k] = cp[k];
            if (l * 64 + k * 16 + 16 > bytes[a]) {  // the tail beyond the array: never compare or send garbage
              uint32_t w[4] = {cur[k].x, cur[k].y, cur[k].z, cur[k].w};
              for (int q = 0; q < 16; ++q)
                if (l * 64 + k * 16 + q >= bytes[a]) w[q >> 2] &= ~(0xFFu << (8 * (q & 3)));
              cur[k] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (differs(cur[k], sh[k])) { sh[k] = cur[k]; changed = true; }
          }
        }
        emit_line(m, changed, (uint32_t)(((dst[a] - m.h_base) >> 6) + l), cur, sent);
      }
      sh_off += lines * 64;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sent += __shfl_xor_sync(RG_FULL, sent, o);
  if (lane == 0 && sent) atomicAdd(m.bytes, (unsigned long long)sent);
  if (publish) {  // the pass that ends the call: the last block to finish hands the counters to the host
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(m.ticket, 1u) == gridDim.x - 1) {
        __threadfence();
        *m.h_bytes = *reinterpret_cast<volatile unsigned long long*>(m.bytes);
        *m.h_errflag = *reinterpret_cast<volatile uint32_t*>(b.errflag);
        *b.errflag = 0u;  // reported: rg_sync clears it the same way
        *m.bytes = 0ull;
        *m.ticket = 0u;
        __threadfence_system();
        const unsigned long long seq = *m.seq + 1ull;
        *m.seq = seq;
        *reinterpret_cast<volatile unsigned long long*>(m.h_seq) = seq;  // everything above is in host memory before this
        __threadfence_system();
      }
    }
  }
}

// ---------------------------------------------------------------- launchers
static int mirror_blocks(const DevBatch& b, int sm_count) {
  return (int)std::min<int64_t>((b.n + 7) / 8, (int64_t)sm_count * 8);  // a warp per env, grid-stride, 64 warps per SM
}
static bool mirror_by_lines(const DevBatch& b, const MirrorArgs& m) { return m.mode != 0 && (b.C & 63) == 0 && b.CP == b.C; }
static const int MIRROR_ITEMS_MAX = 16000;  // words of k_mirror_lines' work list + env table (64 KB of shared memory at most)
// One 1024-thread block per SM: on a few SMs beside the step's other kernels (pass 1, over k_step_fast's list of moves),
// on every SM when the pass runs alone.
static cudaError_t launch_lines(const DevBatch& b, const MirrorArgs& m, int pass, int publish, int sm_count, cudaStream_t s) {
  const int lpe = b.C >> 6, ppe = m.with_hist ? b.HB >> 4 : 0;
  const int threads = 1024;
  const int epb = std::max(32, std::min(threads, MIRROR_ITEMS_MAX / (lpe + ppe + 1)) / 32 * 32);
  const int env_chunks = (int)((b.n + epb - 1) / epb);
  const int blocks = std::max(1, std::min(env_chunks, pass == 1 ? m.confined_sms : sm_count));
  k_mirror_lines<<<blocks, threads, (size_t)epb * (lpe + ppe + 1) * 4, s>>>(b, m, pass, epb, env_chunks, pass != 1, publish, 0, b.n);
  return cudaGetLastError();
}
static size_t one_warp_smem(const DevBatch& b) { return 2 * (size_t)b.CP + sizeof(EnvState) + 16; }
static size_t block_smem(const DevBatch& b) { return (size_t)WARPS_PER_BLOCK * one_warp_smem(b); }
static int pf_warps_per_block(const DevBatch& b) {
  int w = b.pf_wpb > 0 ? b.pf_wpb : PF_MAX_WPB;
  const int fit = (int)((200u << 10) / one_warp_smem(b));  // 227 KB per block at most; leave room for the step kernels
  if (w > fit) w = fit;
  if (w > PF_MAX_WPB) w = PF_MAX_WPB;
  return w < 1 ? 1 : w;
}

cudaError_t configure_kernels(const DevBatch& b) {
  size_t sm = block_smem(b);
  cudaError_t e = cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_step_player, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one_warp_smem(b));

  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_step_monsters, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_step_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GEN_WPB * one_warp_smem(b)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_prefetch, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           b.pf_exclusive ? PF_EXCLUSIVE_SMEM : (int)(pf_warps_per_block(b) * one_warp_smem(b)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_spec_build, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(SPEC_WPB * (one_warp_smem(b) + (size_t)b.CP)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_mirror_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, MIRROR_ITEMS_MAX * 4);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_complete_maps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one_warp_smem(b));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_test_move_enemy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one_warp_smem(b));
}
cudaError_t launch_reset(const DevBatch& b, cudaStream_t s) {
  int blocks = (int)((b.n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  k_reset<<<blocks, WARPS_PER_BLOCK * 32, block_smem(b), s>>>(b);
  return cudaGetLastError();
}
// Enqueues one env-step: the thread-per-env kernel (classification + the cheap steps), the full-path kernel on
// `side` beside the player and monster kernels, the synchronous-reset pass, the join, and the step counter. No
// per-step arguments (the step parity lives on the device, the actions are read from a fixed buffer), so
// the whole sequence is captured once into a CUDA graph and replayed with one launch per step.
cudaError_t launch_step(const DevBatch& b, const uint8_t* actions_src, uint8_t* actions, int auto_reset, const StepStreams& q,
                        const MirrorArgs* mirror, int sm_count) {
  cudaStream_t s = q.main;
  const size_t sm = block_smem(b);
  int gen_blocks = (int)std::min<int64_t>(b.gen_warps / GEN_WPB, (b.n + GEN_WPB - 1) / GEN_WPB);
  const size_t gen_sm = GEN_WPB * one_warp_smem(b);
  const int pblocks = (int)std::min<int64_t>(b.player_blocks > 0 ? b.player_blocks : (int64_t)sm_count * RG_HOT_MIN_BLOCKS, b.n);
  const int mon_blocks = (int)std::min<int64_t>(b.mon_warps / WARPS_PER_BLOCK, (b.n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  cudaError_t e;
  k_step_scan<<<(unsigned)((b.n + 255) / 256), 256, 0, s>>>(b, actions_src, actions);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if ((e = cudaEventRecord(q.ev_fork, s)) != cudaSuccess) return e;
  // branch 1, high-priority side stream: full-path steps (descents, MoveUntil) - a few long serial chains
  if ((e = cudaStreamWaitEvent(q.side, q.ev_fork, 0)) != cudaSuccess) return e;
  k_step_gen<<<gen_blocks, GEN_WPB * 32, gen_sm, q.side>>>(b, actions, auto_reset, 0, 0);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if ((e = cudaEventRecord(q.ev_join, q.side)) != cudaSuccess) return e;
  // branch 2, high-priority stream: the envs with an active monster - player phase, then monster phase.
  // In the host-facing step the mirror's first pass is the longest thing left once k_step_fast has ended (PCIe-bound,
  // ~65 us) and this branch fits beside it: there it starts after k_step_fast, which then has the SMs' registers to
  // itself (31 instead of 44 us) and lets the PCIe writes start that much earlier.
  cudaStream_t m = b.branches ? q.mon : s;
  const bool fast_first = mirror && mirror->fast_first && b.branches;
  auto branch_a = [&](cudaEvent_t after) -> cudaError_t {
    cudaError_t e2;
    if (b.branches && (e2 = cudaStreamWaitEvent(q.mon, after, 0)) != cudaSuccess) return e2;
    k_step_player<<<pblocks, 32, one_warp_smem(b), m>>>(b, actions, auto_reset, 0);
    if ((e2 = cudaGetLastError()) != cudaSuccess) return e2;
    k_step_monsters<<<b.mon_warps / WARPS_PER_BLOCK, WARPS_PER_BLOCK * 32, sm, m>>>(b, auto_reset, 0);
    if ((e2 = cudaGetLastError()) != cudaSuccess) return e2;
    return b.branches ? cudaEventRecord(q.ev_mon, q.mon) : cudaSuccess;
  };
  if (!fast_first && (e = branch_a(q.ev_fork)) != cudaSuccess) return e;
  // branch 3, main stream: every other env, one thread each; then what that kernel could not finish
  k_step_fast<<<(unsigned)((b.n + 127) / 128), 128, 0, s>>>(b, actions, auto_reset);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (mirror && (e = cudaEventRecord(q.ev_fast, s)) != cudaSuccess) return e;
  if (fast_first && (e = branch_a(q.ev_fast)) != cudaSuccess) return e;
  if (mirror) {
    // host mirror, first pass: the envs k_step_fast finished (most of them), beside every other kernel of the
    // step - the pass is dominated by small PCIe writes, not by SM work
    if ((e = cudaStreamWaitEvent(q.mir, q.ev_fast, 0)) != cudaSuccess) return e;
    if (mirror_by_lines(b, *mirror)) {
      if ((e = launch_lines(b, *mirror, 1, 0, sm_count, q.mir)) != cudaSuccess) return e;
    } else {
      k_mirror<<<b.mirror_blocks > 0 ? b.mirror_blocks : mirror_blocks(b, sm_count), 256, 0, q.mir>>>(b, *mirror, 1, 0, b.n);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if ((e = cudaEventRecord(q.ev_mir, q.mir)) != cudaSuccess) return e;
  }
  k_step_player<<<pblocks, 32, one_warp_smem(b), s>>>(b, actions, auto_reset, 1);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  k_step_monsters<<<std::max(1, mon_blocks / 4), WARPS_PER_BLOCK * 32, sm, s>>>(b, auto_reset, 1);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(s, q.ev_join, 0)) != cudaSuccess) return e;
  if (b.branches && (e = cudaStreamWaitEvent(s, q.ev_mon, 0)) != cudaSuccess) return e;
  if (auto_reset) {
    // episode ends whose next game was not prefetched in time (normally none; with prefetching on, a
    // small grid is enough), and the end-of-step bookkeeping in the last block to finish
    const int rblocks = b.prefetch ? std::min(gen_blocks, 296) : gen_blocks;
    k_step_gen<<<rblocks, GEN_WPB * 32, gen_sm, s>>>(b, actions, auto_reset, 1, 1);
  } else {
    k_step_end<<<1, 32, 0, s>>>(b, auto_reset);
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (mirror) {  // second pass: the envs that were finished by the player, monster, full-path and reset kernels
    // (the second pass takes every env that still has rows marked: it must not start before the first has ended; nor
    // may the publishing pass end before the first one's stores have landed)
    if ((e = cudaStreamWaitEvent(s, q.ev_mir, 0)) != cudaSuccess) return e;
    if (mirror_by_lines(b, *mirror)) return launch_lines(b, *mirror, 2, 2, sm_count, s);
    k_mirror<<<mirror_blocks(b, sm_count), 256, 0, s>>>(b, *mirror, 2, 0, b.n);
  }
  return cudaGetLastError();
}
cudaError_t launch_prefetch(const DevBatch& b, int warps, int slot, cudaStream_t s) {
  const int PF_WPB = pf_warps_per_block(b);
  const int64_t blocks_all = (b.n + PF_WPB - 1) / PF_WPB;
  int blocks = (warps > 0 ? warps : b.gen_warps) / PF_WPB;
  if (blocks > blocks_all) blocks = (int)blocks_all;
  if (blocks < 1) blocks = 1;
  // pf_exclusive: the block asks for (nearly) all of an SM's shared memory, so no step-kernel block shares
  // the SM with it - the generator's code and the step kernels' code stop evicting each other
  k_prefetch<<<blocks, PF_WPB * 32, b.pf_exclusive ? (size_t)PF_EXCLUSIVE_SMEM : PF_WPB * one_warp_smem(b), s>>>(b, slot);
  return cudaGetLastError();
}
cudaError_t launch_spec_build(const DevBatch& b, int slot, cudaStream_t s) {
  k_spec_build<<<b.spec_warps / SPEC_WPB, SPEC_WPB * 32, SPEC_WPB * (one_warp_smem(b) + (size_t)b.CP), s>>>(b, slot);
  return cudaGetLastError();
}
cudaError_t launch_test_move_enemy(const DevBatch& b, int64_t env, int fx, int fy, int tx, int ty, int* out3,
                                   cudaStream_t s) {
  k_test_move_enemy<<<1, 32, one_warp_smem(b), s>>>(b, env, fx, fy, tx, ty, out3);
  return cudaGetLastError();
}
cudaError_t launch_encode(const DevBatch& b, int mode, uint32_t flag, int with_hist, int channels, float* out,
                          cudaStream_t s) {
  if (mode == 0 && (b.C & 3) == 0) {
    const int per_env = channels * (b.C / 4);
    dim3 grid((unsigned)((b.n + GRAY_ENVS - 1) / GRAY_ENVS), (unsigned)((per_env + 255) / 256));
    k_encode_gray<<<grid, 256, 0, s>>>(b, flag, channels, out);
  } else {
    k_encode<<<(unsigned)b.n, 256, 0, s>>>(b, mode, flag, with_hist, channels, out);
  }
  return cudaGetLastError();
}
cudaError_t launch_encode_compact(const DevBatch& b, uint8_t* sym_out, int32_t* status_out, uint8_t* hist_out, int sm_count,
                                  cudaStream_t s) {
  const int64_t total = b.n * (b.CP / 16);
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count * 16);
  k_encode_compact<<<blocks, 256, 0, s>>>(b, sym_out, status_out, hist_out);
  return cudaGetLastError();
}
cudaError_t launch_complete_maps(const DevBatch& b, int64_t env_lo, int64_t env_hi, cudaStream_t s) {
  if (env_hi <= env_lo) return cudaSuccess;
  k_complete_maps<<<(unsigned)(env_hi - env_lo), 32, one_warp_smem(b), s>>>(b, env_lo, env_hi);
  return cudaGetLastError();
}
cudaError_t launch_mirror(const DevBatch& b, const MirrorArgs& m, int sm_count, cudaStream_t s) {
  if (mirror_by_lines(b, m)) return launch_lines(b, m, 0, 1, sm_count, s);
  k_mirror<<<mirror_blocks(b, sm_count), 256, 0, s>>>(b, m, 0, 0, b.n);
  return cudaGetLastError();
}
cudaError_t launch_keys_from_index(const DevBatch& b, const void* idx, int index_bytes, uint8_t* keys, cudaStream_t s) {
  k_keys_from_index<<<(unsigned)((b.n + 255) / 256), 256, 0, s>>>(b, idx, index_bytes, keys);
  return cudaGetLastError();
}
cudaError_t launch_train_reward(const DevBatch& b, float stair_reward, int32_t* level_seen, float* reward_out, cudaStream_t s) {
  k_train_reward<<<(unsigned)((b.n + 255) / 256), 256, 0, s>>>(b, stair_reward, level_seen, reward_out);
  return cudaGetLastError();
}
cudaError_t launch_seed(const DevBatch& b, const uint64_t* lo, const uint64_t* hi, int seeded, int64_t count, cudaStream_t s) {
  k_seed<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(b, lo, hi, seeded, count);
  return cudaGetLastError();
}
cudaError_t launch_unpack_hist(const DevBatch& b, uint8_t* out, cudaStream_t s) {
  k_unpack_hist<<<(unsigned)b.n, 128, 0, s>>>(b, out);
  return cudaGetLastError();
}
cudaError_t launch_state_terminal(const DevBatch& b, uint8_t* out, cudaStream_t s) {
  k_state_terminal<<<(unsigned)((b.n + 255) / 256), 256, 0, s>>>(b, out);
  return cudaGetLastError();
}
cudaError_t launch_export_rooms(const DevBatch& b, int16_t* rooms, cudaStream_t s) {
  k_export_rooms<<<(unsigned)((b.n * MAX_ROOMS + 255) / 256), 256, 0, s>>>(b, rooms);
  return cudaGetLastError();
}
cudaError_t launch_state_hash(const DevBatch& b, uint64_t* out, cudaStream_t s) {
  k_state_hash<<<(unsigned)((b.n + 127) / 128), 128, 0, s>>>(b, out);
  return cudaGetLastError();
}

}  // namespace rg
