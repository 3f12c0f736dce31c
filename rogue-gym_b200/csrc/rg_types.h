// rg_types.h — device-resident layout of one batch (shared by the kernels and the host glue).
//
// HBM layout (SoA over envs; every per-env slice is 16-byte aligned so a warp moves it with
// 128-bit loads/stores):
//   surface   u8  [N][CP]      Surface code per cell            (reference Cell.surface, field.rs:12-15)
//   attr      u8  [N][CP]      CellAttr bits 0..5 + bit 6 "in Floor.doors" (field.rs:107-124, floor.rs:19)
//   screen    u8  [N][CP]      composed ASCII screen            (PlayerState.map, python/src/lib.rs:32)
//   hist      u8  [N][HB]      visited bitmap shown to the agent(PlayerState.history, :33)
//   walk      u32 [N][H][WW]   monster-walkable bitboard rows   (derived from surface; feeds the BFS)
//   dist      u16 [N][9][CP]   DistCache maps                   (rogue/mod.rs:492-518)
//   bfs       u32 [N][9][2][H][WW] frontier / visited bitboards of suspended (lazy) BFS maps
//   st        EnvState [N]     everything scalar / small        (RunTime + GameStateImpl + PlayerState.status)
// CP = W*H rounded up to 16, HB = CP/8 rounded up to 16, WW = ceil(W/32).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/rogue_b200.h"

namespace rg {

constexpr int MAX_ROOMS = RG_MAX_ROOMS;
constexpr int NCACHE = RG_DIST_CACHE;
constexpr int TRACE_KERNELS = 12;  // kernels per step slot of the RG_TRACE timeline
constexpr int SP_DEPTH = 2;  // prefetched next-episode games kept per env

// Surface codes follow the reference's declaration order (rogue/mod.rs:137-146).
enum : uint8_t { S_PASSAGE = 0, S_FLOOR = 1, S_WALLX = 2, S_WALLY = 3, S_STAIR = 4, S_DOOR = 5, S_TRAP = 6, S_NONE = 7 };
// CellAttr (field.rs:107-124) + two private bits.
enum : uint8_t {
  A_VISITED = 1, A_HIDDEN = 2, A_VISIBLE = 4, A_DRAWN = 8, A_LOCKED = 16, A_DARK = 32,
  A_DOOR = 64,   // cell is in Floor.doors
  A_MARK = 128   // generation scratch: maze passage member (cleared before the floor is used)
};
enum : uint8_t { K_NORMAL = 0, K_MAZE = 1, K_EMPTY = 2 };
enum : uint8_t { RF_DARK = 1, RF_VISITED = 2, RF_GOLD = 4 };
enum : uint8_t { MF_PRESENT = 1, MF_ACTIVE = 2 };
enum { RGS_SWAP_IN = 0, RGS_SYNC_RESET = 1, RGS_FULL_STEP = 2, RGS_PREFETCH_BUILT = 3, RGS_PREFETCH_STALE = 4,
       RGS_MONSTER_ENVS = 5, RGS_SPEC_HITS = 6, RGS_FAST_STEPS = 7 };
// full_path[env]: which kernel leaves the env's step final (the host mirror's first pass takes the FP_FAST ones)
enum : uint8_t { FP_FAST = 0, FP_FULL = 1, FP_MONSTERS = 2, FP_RESET = 3, FP_PLAYER = 5 };
enum : uint8_t { SF_REDRAW = 1, SF_STATUS = 2, SF_DEAD = 4, SF_SKIP = 8, SF_PANIC = 16 };
// EnemyAttr bits that the path reads (enemies.rs:125-137)
enum : uint32_t { EA_MEAN = 1u, EA_RANDOM = 0x200u, EA_CONFUSED = 0x400u };
// MessageFlagInner (python/src/flags.rs:9-17)
enum : uint32_t {
  MSG_HIT_FROM = 1, MSG_HIT_TO = 2, MSG_MISS_TO = 4, MSG_MISS_FROM = 8, MSG_KILLED = 16, MSG_SECRET_DOOR = 32,
  MSG_NO_DOWNSTAIR = 64
};

struct RoomD {       // rooms.rs:23-41 minus the generation-only cell sets
  uint8_t kind;      // K_*
  uint8_t flags;     // RF_*
  uint8_t x0, y0;    // room / maze range (half-open), or up_left for K_EMPTY
  uint8_t x1, y1;
  uint16_t ncells;   // maze: number of passage cells
};
struct MonD {        // enemies.rs:159-171; level/defense/dice come from the kind table
  uint8_t x, y, kind, flags;
  int32_t hp;
  uint32_t exp;
};

// The first HOT_BYTES bytes hold everything the thread-per-env fast kernel (k_step_fast) reads or writes, as
// 16-byte pieces it loads with 128-bit accesses; the warp-per-env kernels stage the whole struct in shared memory.
struct alignas(16) EnvState {
  // piece 0
  int16_t px, py;
  uint8_t is_terminal, ui_dead, error, seeded;
  uint32_t steps, message;
  // piece 1
  uint32_t food_left, quiet, gold, exp;
  // piece 2
  int32_t level, hp, hp_max, plevel;
  // piece 3: incremental compose - the screen is persistent, so only rows that can differ are recomputed
  uint64_t dirty_rows;     // bit r: a cell of row r changed in the tile planes since the last compose
  uint64_t ov_rows;        // bit r: the last compose drew an overlay (monster, item, player) in row r (may be a superset)
  // piece 4: monster summary (recomputed from mon[] by every write-back of a warp kernel, see write_back)
  uint16_t mon_present;    // bit m: mon[m] is present
  uint16_t mon_active;     // bit m: mon[m] is present and active
  uint32_t episode;        // resets so far (drives fresh seeds when the config has none)
  // per-step hand-over between the phase kernels (player -> monsters -> finish)
  uint32_t f_msg;          // message bits collected so far
  uint32_t f_gold_before;  // displayed gold when the step began (reward = max(0, after - before))
  // pieces 5-6, 7-8
  uint16_t mon_xy[MAX_ROOMS];    // x | y << 8 of mon[m] (valid where mon_present has the bit)
  uint16_t item_pos[MAX_ROOMS];  // y*W+x or 0xFFFF ; slot = room id
  // pieces 9-11
  uint32_t status[10];     // DISPLAYED status (stale semantics, state_impls.rs:63-65)
  uint16_t cache_snap;     // bit s: cache slot s resumes on its private walkability snapshot (wsnap)
  uint8_t f_flags;         // SF_*
  uint8_t cache_n, cache_head;
  uint8_t spec_req;        // 1 = the next level's skeleton has been requested from k_spec_build for this level
  uint16_t stair_pos;      // y*W+x of this level's stair (0xFFFF: none)
  // ---- not touched by the fast kernel
  uint32_t rng[12];        // dungeon, item, enemy xorshift128 states
  uint32_t seed[4];        // seed used by the next reset (thread_impls.rs:125-128)
  uint8_t cache_x[NCACHE + 1], cache_y[NCACHE + 1];
  uint32_t pad3;
  uint16_t cache_lvl[NCACHE + 1];  // BFS levels finished per slot, 0xFFFF = map complete (lazy DistCache)
  uint32_t pad4;
  RoomD rooms[MAX_ROOMS];
  uint32_t item_amt[MAX_ROOMS];
  MonD mon[MAX_ROOMS];           // slot = room id the monster was spawned in
};
constexpr int HOT_BYTES = 192;
static_assert(sizeof(EnvState) % 16 == 0, "EnvState must be a multiple of 16 bytes");
static_assert(offsetof(EnvState, food_left) == 16 && offsetof(EnvState, level) == 32 && offsetof(EnvState, dirty_rows) == 48 &&
                  offsetof(EnvState, mon_present) == 64 && offsetof(EnvState, mon_xy) == 80 &&
                  offsetof(EnvState, item_pos) == 112 && offsetof(EnvState, status) == 144 && offsetof(EnvState, rng) == HOT_BYTES,
              "k_step_fast unpacks the hot block by offset");
static_assert(offsetof(EnvState, spec_req) == 189 && offsetof(EnvState, stair_pos) == 190 && offsetof(EnvState, rooms) % 8 == 0 &&
                  sizeof(RoomD) == 8,
              "layout assumed by k_step_fast / take_spec");

// A level skeleton built ahead of time (k_spec_build): valid for a descent whose dungeon stream and level match.
struct SpecTag {
  uint32_t rd_before[4];   // dungeon stream the skeleton was generated from
  uint32_t rd_after[4];    // ... and where it ended
  int32_t level;           // the level it is the skeleton of
  uint32_t ok;             // 0 = the build hit a panic state: never used
  uint32_t pad[2];
};
static_assert(sizeof(SpecTag) == 48, "take_spec reads the tag with 128-bit loads");
constexpr int SPEC_RADIUS = 6;  // a skeleton is requested when the player comes this close (Chebyshev) to the stair

struct DevBatch {
  int64_t n;
  int32_t W, H, C, CP, HB, WW;
  int32_t nx, ny, rsx, rsy;  // room grid and sector size (rooms.rs:176)
  int32_t gen_warps;         // warps of the (grid-stride) generation kernel
  int64_t max_steps;
  const rg_params* P;  // device copy: one entry, or the table of distinct configs of a heterogeneous batch
  const uint16_t* cfg_idx;  // [N] env -> entry of P, or nullptr when every env uses P[0]
  const uint8_t* room_lut;  // [160] sector column of x, then [48] sector row of y (0xFF = no sector)
  uint8_t* surface;
  uint8_t* attr;
  uint8_t* screen;
  uint8_t* hist;
  uint32_t* walk;
  uint16_t* dist;
  uint32_t* bfs;       // u32 [N][9][2][H][WW] suspended-BFS frontier / visited rows
  uint32_t* wsnap;     // u32 [N][9][H][WW] walkability a suspended map was started on, once the floor has changed
  EnvState* st;
  // observation block
  uint32_t* status;
  int32_t* reward;
  uint8_t* done;
  uint32_t* message;
  uint8_t* error;
  uint32_t* errflag;  // OR of all errors raised since the last rg_sync
  uint64_t* scr_rows; // [N] bit r: row r of screen / history was rewritten since the host mirror last looked
  uint32_t* defer_list;   // [N] env id | DEFER_* : work handed to the full-path kernel k_step_gen
  uint32_t* defer_count;  // [2] ping-pong by step parity
  uint8_t* full_path;     // [N] who finishes this step of the env (FP_*), written by k_step_fast every step
  uint32_t* fast_list_m;  // [N] k_step_fast's envs whose action is a move (k_step_scan groups them: no divergence)
  uint32_t* fast_list_l;  // [N] k_step_fast's other envs
  uint32_t* fast_count;   // [parity][move, other] list lengths; [4] = moves of the step under way (for the host mirror's first pass)
  uint32_t* slow_list;    // [N] envs with an active monster (k_step_scan -> player kernel, first list)
  uint32_t* slow_list_b;  // [N] envs k_step_fast left to the warp-per-env player kernel (second list)
  uint32_t* slow_count;   // [list][parity] lengths, then [list][parity] work cursors of the player kernels
  int32_t mirror_blocks;  // > 0: grid of the first host-mirror pass (RG_MIRROR_BLOCKS)
  int32_t panic_policy;   // 0 sticky (reference-like: the env is dead for good), 1 terminal (see finish_env)
  int32_t branches;       // 1 = the active-monster envs' kernels run on a stream of their own beside k_step_fast
  int32_t fast;           // 0 = k_step_fast only classifies (every env goes to the player kernel; RG_FAST=0)
  uint32_t* reset_list;   // [N] terminal envs whose next game was not prefetched in time (finish_env -> reset pass of k_step_gen)
  uint32_t* reset_count;  // [2]
  // "next episode" buffers, filled in the background by k_prefetch and swapped in by finish_env
  // when an env's episode ends (sp_state: 0 = empty / being built, 1 = ready)
  uint8_t* sp_surface;    // [N][SP_DEPTH][CP]
  uint8_t* sp_attr;       // [N][SP_DEPTH][CP]
  uint8_t* sp_screen;     // [N][SP_DEPTH][CP]
  uint8_t* sp_hist;       // [N][SP_DEPTH][HB]
  uint32_t* sp_walk;      // [N][SP_DEPTH][H][WW]
  EnvState* sp_st;        // [N][SP_DEPTH]  (slot of episode e = e % SP_DEPTH)
  uint8_t* sp_state;      // [N][SP_DEPTH]
  unsigned long long* trace;  // optional (RG_TRACE=1): [512 steps][TRACE_KERNELS][2] first start / last end, globaltimer ns
  int32_t trace_step;         // slot of a background pass (step kernels use *dstep)
  uint32_t* dstep;            // [0] steps executed so far (low bit selects the work-list buffers), [1] auto-reset steps
  unsigned long long* stats;  // [8] RGS_* event counters since creation (observability)
  uint32_t* refill_ring;  // [refill_cap] env ids whose prefetch ring has a free slot (producers: finish_env)
  uint32_t* refill_ctl;   // [0] tail (producers), [2] end of the last window handed to a pass
  uint32_t refill_cap;    // power of two
  uint32_t* refill_win;   // [8][2] window [begin, end) of the ring served by background pass k % 8 (k_step_end)
  uint32_t* sp_lock;      // [N] 1 = a background pass is working on this env's ring (passes overlap)
  uint32_t* sp_cancel;    // [N][SP_DEPTH] episode whose game must not be published: it was built synchronously
  int32_t prefetch_every; // a background pass is kicked every k-th auto-reset step
  int32_t pf_wpb;         // warps per block of k_prefetch (0 = default)
  int32_t pf_exclusive;   // k_prefetch blocks take a whole SM each (see launch_prefetch)
  int32_t prefetch;       // 0 = off (every reset is generated synchronously by k_step_gen)
  // speculative descents: the next level's skeleton per env, built in the background from a snapshot of the
  // dungeon stream, taken by new_level when level and stream still match (seqlock: odd = being written)
  int32_t spec;             // 0 = off (RG_SPEC=0)
  uint8_t* spec_S;          // [N][CP]
  uint8_t* spec_A;          // [N][CP] (A_MARK still set on maze cells)
  RoomD* spec_rooms;        // [N][MAX_ROOMS]
  SpecTag* spec_tag;        // [N]
  uint32_t* spec_seq;       // [N]
  uint32_t* spec_lock;      // [N] 1 = a builder warp is working on this env's slot
  uint32_t* spec_ring;      // [spec_cap] env ids with a pending request (producers: step kernels)
  uint32_t* spec_ctl;       // [0] tail, [2] end of the last window handed to a pass
  uint32_t* spec_win;       // [8][2] window served by the pass kicked after step s % 8
  uint32_t spec_cap;        // power of two
  int32_t spec_warps;
  uint32_t* mon_list;     // [N] envs with an active monster this step (player kernel -> monster kernel), first list
  uint32_t* mon_list_b;   // [N] the same for the second slow list
  uint32_t* mon_count;    // [list][parity] lengths, then [list][parity] work cursors
  int32_t mon_warps;      // warps of the (grid-stride) monster kernel
  int32_t player_blocks;  // > 0: one-warp blocks of the player kernel (default: 32 per SM)
};

}  // namespace rg
