// rg_device.cuh — warp-per-env device code of the B200 Rogue-Gym simulator.
//
// Execution model. One warp owns one dungeon. The env's tile planes (surface, attr) and its
// small state live in shared memory while the warp works on it. Code is written in two modes:
//   * UNIFORM: all 32 lanes execute the same scalar logic on the same data (decisions, RNG
//     draws with their rejection loops, entity bookkeeping). Every lane computes the same
//     value, so there is no divergence and no broadcast traffic; stores to shared memory are
//     same-address/same-value.
//   * PARALLEL: lanes take different cells / rows / directions / entities (tile fills, room
//     reveal, bit-parallel BFS with one bitboard row per lane, the 9-neighbour probe of a
//     monster, screen compose with 128-bit accesses). __syncwarp() separates a PARALLEL write
//     phase from whatever reads it next.
// This is a new design, not a translation: the reference's FenwickSets become popcount/nth
// selects over implicit sets, its BTreeMaps a 16-slot monster table ordered on the fly, its
// queue BFS a level-synchronous bitboard sweep, and the deferred passage list a two-pass RNG
// replay. What must stay identical is the observable result and the order of RNG draws; each
// function cites the reference code whose results it reproduces (/root/reference, c78608b).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rg_types.h"

namespace rg {

#define RG_FULL 0xffffffffu
// Every kernel stages its env in the one dynamic shared array. Device functions re-derive the
// plane / state pointers from this symbol (RG_PLANES) instead of carrying generic pointers, so
// that accesses compile to LDS/STS even across non-inlined calls.
extern __shared__ __align__(16) unsigned char rg_smem[];
#define RG_PLANES(c)                                                          \
  uint8_t* const S = rg_smem + (c).soff;                                      \
  uint8_t* const A = S + (c).CP;                                              \
  EnvState* const st = reinterpret_cast<EnvState*>(A + (c).CP);               \
  (void)S; (void)A; (void)st
#define RG_DEV __device__ __forceinline__

// Direction order Up, Down, Left, Right, LeftUp, RightUp, LeftDown, RightDown, Stay (coord.rs:198-208)
enum { D_UP = 0, D_DOWN, D_LEFT, D_RIGHT, D_LEFTUP, D_RIGHTUP, D_LEFTDOWN, D_RIGHTDOWN, D_STAY };
RG_DEV int ddx(int d) { return (int)((0x18885u >> (2 * d)) & 3u) - 1; }
RG_DEV int ddy(int d) { return (int)((0x1A058u >> (2 * d)) & 3u) - 1; }
RG_DEV bool is_diag(int d) { return d >= D_LEFTUP && d <= D_RIGHTDOWN; }
RG_DEV int reverse_dir(int d) { return (int)((0x845672301ull >> (4 * d)) & 15ull); }  // coord.rs:272-285
RG_DEV bool can_walk(uint8_t s) { return !(s == S_WALLX || s == S_WALLY || s == S_NONE); }  // rogue/mod.rs:176-183
RG_DEV uint8_t surface_tile(uint8_t s) { return (uint8_t)(0x205E2B257C2D2E23ull >> (8 * s)); }  // rogue/mod.rs:148-161

// ------------------------------------------------------------------ RNG
// xorshift128 (rand_xorshift 0.2) + rand 0.7 UniformInt::sample_single; reference wrapper: rng.rs:47-98
struct Rng {
  uint32_t x, y, z, w;
  RG_DEV void load(const uint32_t* p) { x = p[0]; y = p[1]; z = p[2]; w = p[3]; }
  RG_DEV void store(uint32_t* p) const { p[0] = x; p[1] = y; p[2] = z; p[3] = w; }
  RG_DEV void seed(const uint32_t* s) {
    x = s[0]; y = s[1]; z = s[2]; w = s[3];
    if ((x | y | z | w) == 0) x = y = z = w = 0x0BAD5EEDu;
  }
  RG_DEV uint32_t next() {
    uint32_t t = x ^ (x << 11);
    x = y; y = z; z = w;
    w = w ^ (w >> 19) ^ (t ^ (t >> 8));
    return w;
  }
  RG_DEV uint64_t next64() {
    uint64_t lo = next();
    uint64_t hi = next();
    return (hi << 32) | lo;
  }
  // 32-bit lane: u32 and i32 call sites (lo/hi are wrapping)
  RG_DEV uint32_t range32(uint32_t lo, uint32_t hi) {
    uint32_t range = hi - lo;
    uint32_t zone = (range << __clz(range)) - 1u;
    for (;;) {
      uint32_t v = next();
      uint32_t l = v * range;
      if (l <= zone) return lo + __umulhi(v, range);
    }
  }
  RG_DEV int range_i32(int lo, int hi) { return (int)range32((uint32_t)lo, (uint32_t)hi); }
  // 64-bit lane: usize / i64 call sites
  RG_DEV uint64_t range64(uint64_t lo, uint64_t hi) {
    uint64_t range = hi - lo;
    uint64_t zone = (range << __clzll(range)) - 1ull;
    for (;;) {
      uint64_t v = next64();
      uint64_t l = v * range;
      if (l <= zone) return lo + __umul64hi(v, range);
    }
  }
  RG_DEV bool does_happen(uint32_t p_inv) { return range32(0, p_inv) == 0; }  // rng.rs:91-93
  RG_DEV bool parcent(uint32_t p) { return range32(1, 101) <= p; }            // rng.rs:95-98
  // The same samplers as real calls, for the floor generator (the ...g names): it draws at ~100 places,
  // and one shared copy that stays in the instruction caches beats 100 inlined ones that each have to
  // be fetched (measured: reset-only throughput 6.1 M -> 8.6 M floors/s at 80x24).
  __device__ __noinline__ uint32_t range32g(uint32_t lo, uint32_t hi) { return range32(lo, hi); }
  __device__ __noinline__ uint64_t range64g(uint64_t lo, uint64_t hi) { return range64(lo, hi); }
  RG_DEV int range_i32g(int lo, int hi) { return (int)range32g((uint32_t)lo, (uint32_t)hi); }
  RG_DEV bool does_happeng(uint32_t p_inv) { return range32g(0, p_inv) == 0; }
  RG_DEV bool parcentg(uint32_t p) { return range32g(1, 101) <= p; }
};

// Position of the n-th (0-based) set bit of m, which must exist: five popcount halvings instead of
// clearing n bits one after the other (n reaches 31 in the ballot-based cell selects, and this sits on
// the serial chain of a floor).
RG_DEV int nth_set_bit(uint32_t m, uint32_t n) {
  int pos = 0;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const uint32_t c = (uint32_t)__popc(m & (((1u << w) - 1u) << pos));
    if (n >= c) {
      n -= c;
      pos += w;
    }
  }
  return pos;
}

// ------------------------------------------------------------------ per-warp context
struct Ctx {
  uint32_t soff;      // byte offset of this warp's region in rg_smem: surface [CP], attr [CP], EnvState
  uint8_t* S;         // the same three as pointers (kernel-level code only)
  uint8_t* A;
  EnvState* st;
  const uint8_t* col_room;  // global [160] sector column of x (0xFF = none)
  const uint8_t* row_room;  // global [48]  sector row of y (0xFF = none: rows 0, H-1 and beyond the grid)
  const rg_params* P; // global (read-only)
  int W, H, C, CP, WW;
  int lane;
  int nx, ny, rsx, rsy, nrooms;
  uint8_t* g_screen;  // this env's slices in HBM
  uint8_t* g_hist;
  uint64_t* g_rows;   // this env's "screen rows rewritten" mask for the host mirror (nullptr: not the live screen)
  uint32_t* g_walk;
  uint16_t* g_dist;
  uint32_t* g_bfs;    // per cache slot: frontier rows then visited rows of a suspended BFS
  uint32_t* g_wsnap;  // per cache slot: private walkability snapshot (see snapshot_suspended_maps)
  // this env's speculative next-level skeleton (nullptr = off); see take_spec
  const uint8_t* g_spec_S;
  const uint8_t* g_spec_A;
  const RoomD* g_spec_rooms;
  const SpecTag* g_spec_tag;
  const uint32_t* g_spec_seq;
  unsigned long long* g_spec_hits;
  Rng rd, ri, re;     // dungeon / item / enemy streams (registers)
  // per-step reaction summary (state_impls.rs:57-75 collapses to these)
  uint32_t redraw, status_upd, dead, msg, hist_done, a_dirty, s_dirty, panic;
};

RG_DEV void set_panic(Ctx& c) { c.panic = 1; }
// rows [y0, y1] of the tile planes changed: the next compose has to recompute them
RG_DEV void mark_rows(Ctx& c, int y0, int y1) {
  y0 = max(y0, 0);
  y1 = min(y1, c.H - 1);
  if (y1 < y0) return;
  c.st->dirty_rows |= ((~0ull) >> (63 - (y1 - y0))) << y0;
}

// Room grid geometry: rooms.rs:176,191-206
RG_DEV void room_area(const Ctx& c, int i, int& ax0, int& ay0, int& ax1, int& ay1) {
  int xi = i % c.nx, yi = i / c.nx;
  int ry = c.rsy;
  if (yi == 0) {
    ry -= 1;
    ay0 = 1;
  } else {
    ay0 = ry * yi;
  }
  ax0 = c.rsx * xi;
  if (ay0 + ry == c.H) ry -= 1;
  ax1 = ax0 + c.rsx;
  ay1 = ay0 + ry;
}
// Floor::cd_to_room_id floor.rs:194-200 (areas are disjoint, so "first" = "the"). The sector of a
// column / row is a pure function of the config; it is tabulated once per batch by the host (rg_api.cu).
RG_DEV int room_of(const Ctx& c, int x, int y) {
  if (x < 0 || y < 0 || x >= c.W || y >= c.H) return -1;
  const uint32_t cx = __ldg(c.col_room + x), ry = __ldg(c.row_room + y);
  return (cx == 0xFFu || ry == 0xFFu) ? -1 : (int)(ry * c.nx + cx);
}
RG_DEV bool in_rect(const RoomD& r, int x, int y) { return x >= r.x0 && x < r.x1 && y >= r.y0 && y < r.y1; }
RG_DEV bool inb(const Ctx& c, int x, int y) { return x >= 0 && y >= 0 && x < c.W && y < c.H; }
RG_DEV uint32_t lev_add(const Ctx& c) {  // rogue/mod.rs:483-489
  RG_PLANES(c);
  uint32_t lv = (uint32_t)st->level;
  return c.P->amulet_level < lv ? lv - c.P->amulet_level : 0u;
}

// Floor::can_move_impl floor.rs:169-182
RG_DEV bool can_move(const Ctx& c, int x, int y, int d, bool is_enemy) {
  RG_PLANES(c);
  int nx = x + ddx(d), ny = y + ddy(d);
  if (!inb(c, nx, ny)) return false;
  int ni = ny * c.W + nx;
  bool res = can_walk(S[ni]);
  if (!is_enemy) res = res && !(A[ni] & (A_HIDDEN | A_LOCKED));
  if (is_diag(d)) {
    res = res && can_walk(S[y * c.W + nx]);
    res = res && can_walk(S[ny * c.W + x]);
  }
  return res;
}

// ------------------------------------------------------------------ floor generation (rg_gen.inl, two variants)
namespace gen_inl {
#define range32G range32
#define range64G range64
#define range_i32G range_i32
#define does_happenG does_happen
#define parcentG parcent
#include "rg_gen.inl"
#undef range32G
#undef range64G
#undef range_i32G
#undef does_happenG
#undef parcentG
}  // namespace gen_inl
namespace gen_call {
#define range32G range32g
#define range64G range64g
#define range_i32G range_i32g
#define does_happenG does_happeng
#define parcentG parcentg
#include "rg_gen.inl"
#undef range32G
#undef range64G
#undef range_i32G
#undef does_happenG
#undef parcentG
}  // namespace gen_call

// ------------------------------------------------------------------ visibility ("FOV")
// Floor::enters_room floor.rs:231-247 (+ activate_area from player_in :273-279)
__device__ void enters_room(Ctx& c, int x, int y) {
  RG_PLANES(c);
  int id = room_of(c, x, y);
  if (id < 0) { set_panic(c); return; }
  RoomD rm = st->rooms[id];
  if (!(rm.flags & RF_VISITED)) {
    st->rooms[id].flags = rm.flags | RF_VISITED;
    if (rm.kind == K_NORMAL && !(rm.flags & RF_DARK)) {
      int w = rm.x1 - rm.x0, n = w * (rm.y1 - rm.y0);
      for (int k = c.lane; k < n; k += 32) A[(rm.y0 + k / w) * c.W + rm.x0 + k % w] |= (A_DRAWN | A_VISIBLE);
      mark_rows(c, rm.y0, rm.y1 - 1);
      __syncwarp();
    }
  }
  // EnemyHandler::activate_area enemies.rs:342-355: placed MEAN monsters inside the assigned area
  for (int m = 0; m < c.nrooms; ++m) {
    MonD mo = st->mon[m];
    if ((mo.flags & MF_PRESENT) && !(mo.flags & MF_ACTIVE) && (c.P->enemies[mo.kind].attr & EA_MEAN) &&
        room_of(c, mo.x, mo.y) == id)
      st->mon[m].flags = mo.flags | MF_ACTIVE;
  }
}
// Floor::leaves_room floor.rs:250-261
__device__ void leaves_room(Ctx& c, int x, int y) {
  RG_PLANES(c);
  int id = room_of(c, x, y);
  if (id < 0) { set_panic(c); return; }
  RoomD rm = st->rooms[id];
  if (!((rm.flags & RF_VISITED) && (rm.flags & RF_DARK))) return;
  int x0 = rm.x0, y0 = rm.y0, x1 = rm.x1, y1 = rm.y1;
  if (rm.kind == K_EMPTY) room_area(c, id, x0, y0, x1, y1);
  int w = x1 - x0 - 2, h = y1 - y0 - 2;
  if (w <= 0 || h <= 0) return;
  for (int k = c.lane; k < w * h; k += 32) A[(y0 + 1 + k / w) * c.W + x0 + 1 + k % w] &= (uint8_t)~A_VISIBLE;
  mark_rows(c, y0 + 1, y1 - 2);
  __syncwarp();
}
// Floor::player_in floor.rs:264-295 ; Cell::approached field.rs:20-26
__device__ void player_in(Ctx& c, int x, int y, bool init) {
  RG_PLANES(c);
  const int W = c.W;
  if (init || (A[y * W + x] & A_DOOR)) enters_room(c, x, y);
  mark_rows(c, y - 1, y + 1);
  A[y * W + x] |= A_VISITED;
#pragma unroll
  for (int d = 0; d < 9; ++d) {
    int nx = x + ddx(d), ny = y + ddy(d);
    if (!inb(c, nx, ny)) continue;
    int idx = ny * W + nx;
    if (is_diag(d) && S[idx] == S_PASSAGE) continue;
    uint8_t a = A[idx];
    if (a & A_HIDDEN) continue;
    A[idx] = a | A_DRAWN | A_VISIBLE;
  }
  c.a_dirty = 1;
}
// Floor::player_out floor.rs:298-312 ; Cell::left field.rs:30-34
__device__ void player_out(Ctx& c, int x, int y) {
  RG_PLANES(c);
  const int W = c.W;
  if (A[y * W + x] & A_DOOR) leaves_room(c, x, y);
  mark_rows(c, y - 1, y + 1);
#pragma unroll
  for (int d = 0; d < 9; ++d) {
    int nx = x + ddx(d), ny = y + ddy(d);
    if (!inb(c, nx, ny)) continue;
    int idx = ny * W + nx;
    if (S[idx] == S_FLOOR && (A[idx] & A_DARK)) A[idx] &= (uint8_t)~A_VISIBLE;
  }
  c.a_dirty = 1;
}

// Build the monster-walkable bitboard rows from the surface plane (one ballot per 32 cells).
__device__ void build_walk(Ctx& c) {
  RG_PLANES(c);
  __syncwarp();
  for (int row = 0; row < c.H; ++row)
    for (int w = 0; w < c.WW; ++w) {
      int x = w * 32 + c.lane;
      bool b = x < c.W && can_walk(S[row * c.W + x]);
      uint32_t m = __ballot_sync(RG_FULL, b);
      if (c.lane == 0) c.g_walk[row * c.WW + w] = m;
    }
  __syncwarp();
}

__device__ void snapshot_suspended_maps(Ctx& c);  // lazy DistCache, defined with the BFS below

// A descent looks for the skeleton of the level it is about to generate (rooms, mazes, passages, attributes:
// gen_skeleton) among the ones k_spec_build produced in the background. It is taken only if it was generated for
// exactly this level from exactly the dungeon-stream state the descent starts with - then it is, bit for bit,
// what gen_skeleton would produce now - and if the slot was not rewritten while it was being copied (seqlock).
// Returns false with the planes in an undefined state; the caller then generates (which overwrites everything).
__device__ bool take_spec(Ctx& c, Rng& rd, uint32_t level) {
  if (!c.g_spec_seq) return false;
  RG_PLANES(c);
  const uint32_t s1 = *reinterpret_cast<const volatile uint32_t*>(c.g_spec_seq);
  __threadfence();
  const uint4 before = __ldcg(reinterpret_cast<const uint4*>(c.g_spec_tag->rd_before));
  const uint4 after = __ldcg(reinterpret_cast<const uint4*>(c.g_spec_tag->rd_after));
  const int2 lv_ok = __ldcg(reinterpret_cast<const int2*>(&c.g_spec_tag->level));
  bool good = !(s1 & 1u) && lv_ok.y != 0 && lv_ok.x == (int)level && before.x == rd.x && before.y == rd.y &&
              before.z == rd.z && before.w == rd.w;
  good = __shfl_sync(RG_FULL, good ? 1 : 0, 0) != 0;  // one decision per warp
  if (!good) return false;
  __syncwarp();
  for (int i = c.lane; i < c.CP / 16; i += 32) {
    reinterpret_cast<uint4*>(S)[i] = __ldcg(reinterpret_cast<const uint4*>(c.g_spec_S) + i);
    reinterpret_cast<uint4*>(A)[i] = __ldcg(reinterpret_cast<const uint4*>(c.g_spec_A) + i);
  }
  if (c.lane < MAX_ROOMS) {
    const uint2 r = __ldcg(reinterpret_cast<const uint2*>(c.g_spec_rooms) + c.lane);
    reinterpret_cast<uint2*>(st->rooms)[c.lane] = r;
  }
  __threadfence();
  __syncwarp();
  uint32_t same = *reinterpret_cast<const volatile uint32_t*>(c.g_spec_seq) == s1 ? 1u : 0u;
  same = __shfl_sync(RG_FULL, same, 0);
  if (!same) return false;
  rd.x = after.x; rd.y = after.y; rd.z = after.z; rd.w = after.w;
  if (c.lane == 0 && c.g_spec_hits) atomicAdd(c.g_spec_hits, 1ull);
  return true;
}

#define RG_GEN_SECOND_HALF
namespace gen_inl {
#define RG_GEN_NEW_LEVEL_ATTR __device__ __forceinline__
#define range32G range32
#define range64G range64
#define range_i32G range_i32
#define does_happenG does_happen
#define parcentG parcent
#include "rg_gen.inl"
#undef range32G
#undef range64G
#undef range_i32G
#undef does_happenG
#undef parcentG
#undef RG_GEN_NEW_LEVEL_ATTR
}  // namespace gen_inl
namespace gen_call {
#define RG_GEN_NEW_LEVEL_ATTR __device__ __noinline__
#define range32G range32g
#define range64G range64g
#define range_i32G range_i32g
#define does_happenG does_happeng
#define parcentG parcentg
#include "rg_gen.inl"
#undef range32G
#undef range64G
#undef range_i32G
#undef does_happenG
#undef parcentG
#undef RG_GEN_NEW_LEVEL_ATTR
}  // namespace gen_call
#undef RG_GEN_SECOND_HALF

// ------------------------------------------------------------------ BFS distance map
// Floor::make_dist_map floor.rs:395-416 (8-dir, monster rules, diagonal needs both orthogonal
// neighbours walkable) as a level-synchronous bitboard sweep: lane r owns row r (and r+32).
template <int WORDS>
RG_DEV void shl1(const uint32_t (&a)[WORDS], uint32_t (&o)[WORDS]) {
#pragma unroll
  for (int w = WORDS - 1; w > 0; --w) o[w] = (a[w] << 1) | (a[w - 1] >> 31);
  o[0] = a[0] << 1;
}
template <int WORDS>
RG_DEV void shr1(const uint32_t (&a)[WORDS], uint32_t (&o)[WORDS]) {
#pragma unroll
  for (int w = 0; w < WORDS - 1; ++w) o[w] = (a[w] >> 1) | (a[w + 1] << 31);
  o[WORDS - 1] = a[WORDS - 1] >> 1;
}
template <int WORDS, int RPL>
RG_DEV void rows_up(const uint32_t (&X)[RPL][WORDS], uint32_t (&O)[RPL][WORDS], int lane) {  // O[row] = X[row-1]
#pragma unroll
  for (int j = 0; j < RPL; ++j)
#pragma unroll
    for (int w = 0; w < WORDS; ++w) {
      uint32_t same = __shfl_up_sync(RG_FULL, X[j][w], 1);
      uint32_t wrap = (j > 0) ? __shfl_sync(RG_FULL, X[j > 0 ? j - 1 : 0][w], 31) : 0u;
      O[j][w] = lane > 0 ? same : wrap;
    }
}
template <int WORDS, int RPL>
RG_DEV void rows_down(const uint32_t (&X)[RPL][WORDS], uint32_t (&O)[RPL][WORDS], int lane) {  // O[row] = X[row+1]
#pragma unroll
  for (int j = 0; j < RPL; ++j)
#pragma unroll
    for (int w = 0; w < WORDS; ++w) {
      uint32_t same = __shfl_down_sync(RG_FULL, X[j][w], 1);
      uint32_t wrap = (j + 1 < RPL) ? __shfl_sync(RG_FULL, X[j + 1 < RPL ? j + 1 : j][w], 0) : 0u;
      O[j][w] = lane < 31 ? same : wrap;
    }
}

// Lazy evaluation. The reference computes the whole map when a DistCache entry is created
// (rogue/mod.rs:504-517) but only ever reads the nine cells around a monster (:356-368). Here a
// map is extended level by level only until every walkable cell of the 3x3 block that is about
// to be read has its final distance (a BFS label never changes once written), then suspended:
// the frontier / visited bitboards and the number of finished levels are kept per cache slot and
// a later reader resumes from there. Every value that is read is therefore identical to the
// reference's; `complete_all_maps` finishes all suspended maps before the walkability they were
// started on changes (search unlocking a cell, descending to a new floor).
constexpr uint16_t BFS_COMPLETE = 0xFFFFu;

template <int WORDS, int RPL>
__device__ __noinline__ uint32_t bfs_impl(const uint32_t* __restrict__ walk, uint16_t* __restrict__ out,
                                          uint32_t* __restrict__ Fg, uint32_t* __restrict__ Vg, int W, int H, int CP,
                                          int fx, int fy, uint32_t done_levels, int nx0, int ny0, int lane) {
  uint32_t Wc[RPL][WORDS], Wu[RPL][WORDS], Wd[RPL][WORDS], Vis[RPL][WORDS], F[RPL][WORDS], Need[RPL][WORDS];
  const bool fresh = done_levels == 0;
  const bool need_all = nx0 < -1;  // complete the map
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    const int row = lane + 32 * j;
#pragma unroll
    for (int w = 0; w < WORDS; ++w) {
      Wc[j][w] = row < H ? walk[row * WORDS + w] : 0u;
      if (fresh) {
        F[j][w] = (row == fy && (fx >> 5) == w) ? (1u << (fx & 31)) : 0u;
        Vis[j][w] = F[j][w];
      } else {
        F[j][w] = row < H ? Fg[row * WORDS + w] : 0u;
        Vis[j][w] = row < H ? Vg[row * WORDS + w] : 0u;
      }
      // the cells about to be read: columns nx0-1..nx0+1 of rows ny0-1..ny0+1, walkable ones only
      uint32_t nd = 0;
      if (need_all) {
        nd = Wc[j][w];
      } else if (row >= ny0 - 1 && row <= ny0 + 1) {
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = nx0 + dx;
          if (x >= 0 && x < W && (x >> 5) == w) nd |= 1u << (x & 31);
        }
        nd &= Wc[j][w];
      }
      Need[j][w] = nd;
    }
  }
  rows_up<WORDS, RPL>(Wc, Wu, lane);
  rows_down<WORDS, RPL>(Wc, Wd, lane);
  if (fresh) {
    for (int i = lane; i < CP / 8; i += 32)
      reinterpret_cast<uint4*>(out)[i] = make_uint4(RG_FULL, RG_FULL, RG_FULL, RG_FULL);
    __syncwarp();
    if (lane == (fy & 31)) out[fy * W + fx] = 0;
  }
  uint32_t level = done_levels;
  bool complete = false;
  for (;;) {
    uint32_t unmet = 0;
#pragma unroll
    for (int j = 0; j < RPL; ++j)
#pragma unroll
      for (int w = 0; w < WORDS; ++w) unmet |= Need[j][w] & ~Vis[j][w];
    if (!__any_sync(RG_FULL, unmet != 0)) {
      complete = need_all;  // every walkable cell is labelled: nothing can ever be added
      break;
    }
    ++level;
    uint32_t Fu[RPL][WORDS], Fd[RPL][WORDS];
    rows_up<WORDS, RPL>(F, Fu, lane);
    rows_down<WORDS, RPL>(F, Fd, lane);
    uint32_t anynew = 0;
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
      uint32_t a[WORDS], b[WORDS], t0[WORDS], t1[WORDS], n[WORDS];
      shl1<WORDS>(F[j], t0);
      shr1<WORDS>(F[j], t1);
#pragma unroll
      for (int w = 0; w < WORDS; ++w) {
        n[w] = t0[w] | t1[w] | Fu[j][w] | Fd[j][w];
        a[w] = Fu[j][w] & Wc[j][w];  // sources one row up whose vertical step is legal
        b[w] = Fd[j][w] & Wc[j][w];
      }
      shl1<WORDS>(a, t0);
      shr1<WORDS>(a, t1);
#pragma unroll
      for (int w = 0; w < WORDS; ++w) n[w] |= (t0[w] | t1[w]) & Wu[j][w];
      shl1<WORDS>(b, t0);
      shr1<WORDS>(b, t1);
#pragma unroll
      for (int w = 0; w < WORDS; ++w) {
        n[w] |= (t0[w] | t1[w]) & Wd[j][w];
        n[w] &= Wc[j][w] & ~Vis[j][w];
        Vis[j][w] |= n[w];
        F[j][w] = n[w];
        anynew |= n[w];
        uint32_t bits = n[w];
        const int base = (lane + 32 * j) * W + w * 32;
        while (bits) {
          const int bpos = __ffs(bits) - 1;
          bits &= bits - 1;
          out[base + bpos] = (uint16_t)level;
        }
      }
    }
    if (!__any_sync(RG_FULL, anynew != 0) || level >= 0xFFFEu) {
      complete = true;
      break;
    }
  }
  if (!complete) {
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
      const int row = lane + 32 * j;
      if (row < H) {
#pragma unroll
        for (int w = 0; w < WORDS; ++w) {
          Fg[row * WORDS + w] = F[j][w];
          Vg[row * WORDS + w] = Vis[j][w];
        }
      }
    }
  }
  __syncwarp();
  return complete ? (uint32_t)BFS_COMPLETE : level;
}

// Extends the map in `slot` until the 3x3 block around (nx0, ny0) is final (nx0 = -2: until complete).
__device__ void bfs_extend(Ctx& c, int slot, int nx0, int ny0) {
  RG_PLANES(c);
  const uint32_t lvl = st->cache_lvl[slot];
  if (lvl == BFS_COMPLETE) return;
  uint16_t* out = c.g_dist + (size_t)slot * c.CP;
  uint32_t* Fg = c.g_bfs + (size_t)slot * 2 * c.H * c.WW;
  uint32_t* Vg = Fg + c.H * c.WW;
  const int fx = st->cache_x[slot], fy = st->cache_y[slot];
  const int rpl = (c.H + 31) / 32;
  const uint32_t* walk = ((st->cache_snap >> slot) & 1u) ? c.g_wsnap + (size_t)slot * c.H * c.WW : c.g_walk;
  uint32_t r = BFS_COMPLETE;
  bool ok = false;
#define RG_BFS_CASE(WD, RP)                                                                                   \
  if (c.WW == WD && rpl == RP) {                                                                              \
    r = bfs_impl<WD, RP>(walk, out, Fg, Vg, c.W, c.H, c.CP, fx, fy, lvl, nx0, ny0, c.lane);                      \
    ok = true;                                                                                                \
  }
  RG_BFS_CASE(3, 1) else RG_BFS_CASE(1, 1) else RG_BFS_CASE(2, 1) else RG_BFS_CASE(4, 1) else RG_BFS_CASE(5, 1)
  else RG_BFS_CASE(1, 2) else RG_BFS_CASE(2, 2) else RG_BFS_CASE(3, 2) else RG_BFS_CASE(4, 2) else RG_BFS_CASE(5, 2)
#undef RG_BFS_CASE
  if (!ok) {
    set_panic(c);
    return;
  }
  st->cache_lvl[slot] = (uint16_t)r;
  __syncwarp();
}

// Finish every suspended map of this env (parity harness: rg_dump_env shows whole maps).
__device__ void complete_all_maps(Ctx& c) {
  RG_PLANES(c);
  const int n = st->cache_n, head = st->cache_head;
  for (int k = 0; k < n; ++k) {
    const int slot = (head + k) % NCACHE;
    if (st->cache_lvl[slot] != BFS_COMPLETE) bfs_extend(c, slot, -2, -2);
  }
}

// Called before the walkability changes (search unlocking a cell, descending): every suspended
// map that still resumes on the shared bitboard gets a private copy of it, so that whenever it is
// extended later it sees the floor it was started on - exactly what the reference's eagerly
// computed map saw. (Copying 288 B per suspended map replaces finishing up to nine whole BFS.)
__device__ void snapshot_suspended_maps(Ctx& c) {
  RG_PLANES(c);
  const int n = st->cache_n, head = st->cache_head;
  const int words = c.H * c.WW;
  uint32_t snap = st->cache_snap;
  for (int k = 0; k < n; ++k) {
    const int slot = (head + k) % NCACHE;
    if (st->cache_lvl[slot] == BFS_COMPLETE || ((snap >> slot) & 1u)) continue;
    uint32_t* dst = c.g_wsnap + (size_t)slot * words;
    for (int i = c.lane; i < words; i += 32) dst[i] = c.g_walk[i];
    snap |= 1u << slot;
  }
  __syncwarp();
  st->cache_snap = (uint16_t)snap;
  __syncwarp();
}

// DistCache::make_dist_map rogue/mod.rs:504-517: FIFO of 9 keyed by the target coordinate,
// never flushed (stale maps from earlier floors are used on purpose).
// Returns the slot; the map is final at least on the 3x3 block around (nx0, ny0).
__device__ int cached_dist_map(Ctx& c, int tx, int ty, int nx0, int ny0) {
  RG_PLANES(c);
  int n = st->cache_n, head = st->cache_head;
  int slot = -1;
  for (int k = 0; k < n; ++k) {
    int s = (head + k) % NCACHE;
    if (slot < 0 && st->cache_x[s] == tx && st->cache_y[s] == ty) slot = s;
  }
  if (slot < 0) {
    if (n < NCACHE) {
      slot = (head + n) % NCACHE;
      st->cache_n = (uint8_t)(n + 1);
    } else {
      slot = head;
      st->cache_head = (uint8_t)((head + 1) % NCACHE);
    }
    st->cache_x[slot] = (uint8_t)tx;
    st->cache_y[slot] = (uint8_t)ty;
    st->cache_lvl[slot] = 0;
    st->cache_snap = (uint16_t)(st->cache_snap & ~(1u << slot));  // a new map starts on the current floor
    __syncwarp();
  }
  if (nx0 > -100) bfs_extend(c, slot, nx0, ny0);  // nx0 <= -100: only find / create the entry
  return slot;
}

// ------------------------------------------------------------------ monsters
// `moved` = monsters already re-inserted this turn; skip(q) = a moved-active or placed monster
// stands on q (enemies.rs:383-384).
__device__ __noinline__ bool cell_blocked(const Ctx& c, int x, int y, uint32_t moved) {
  RG_PLANES(c);
  bool hit = false;
#pragma unroll 1
  for (int m = 0; m < c.nrooms; ++m) {
    MonD mo = st->mon[m];
    bool counts = (mo.flags & MF_PRESENT) && (!(mo.flags & MF_ACTIVE) || ((moved >> m) & 1u));
    hit = hit || (counts && mo.x == x && mo.y == y);
  }
  return hit;
}

enum { MV_CANT = 0, MV_CAN = 1, MV_REACH = 2 };
// rogue::Dungeon::move_enemy rogue/mod.rs:339-375. PARALLEL over the 9 directions: lane d
// probes neighbour d, the in-order semantics are rebuilt from ballots.
RG_DEV bool room_interior(const RoomD& r, int x, int y) { return x > r.x0 && x < r.x1 - 1 && y > r.y0 && y < r.y1 - 1; }

__device__ int move_enemy(Ctx& c, int mx, int my, int tx, int ty, uint32_t moved, bool noskip, int& ox, int& oy) {
  const int slot = cached_dist_map(c, tx, ty, -100, -100);  // the DistCache entry (FIFO order is observable); no BFS yet
  const uint16_t* dm = c.g_dist + (size_t)slot * c.CP;
  const int d = c.lane;
  const bool valid = d < 9;
  const int nx = mx + ddx(valid ? d : 8), ny = my + ddy(valid ? d : 8);
  // skip(q) for the nine probed cells at once: lane m looks at monster m and, if it counts (placed, or active and
  // already moved: enemies.rs:383-384) and stands inside the 3x3 block, sets the bit of the direction that leads to it
  bool skip = false;
  if (!noskip) {
    RG_PLANES(c);
    uint32_t bit = 0;
    if (c.lane < c.nrooms) {
      const MonD mo = st->mon[c.lane];
      const bool counts = (mo.flags & MF_PRESENT) && (!(mo.flags & MF_ACTIVE) || ((moved >> c.lane) & 1u));
      const int dx = (int)mo.x - mx, dy = (int)mo.y - my;
      if (counts && dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1)  // (dy+1)*3 + (dx+1) -> direction index (coord.rs:198-208)
        bit = 1u << (uint32_t)((0x716382504ull >> (4 * ((dy + 1) * 3 + dx + 1))) & 15ull);
    }
    const uint32_t occ = __reduce_or_sync(RG_FULL, bit);
    skip = valid && ((occ >> d) & 1u);
  }
  const bool live = valid && !skip;
  const bool oob = live && !inb(c, nx, ny);
  uint32_t nd = 0xFFFFu;
  // The common chase happens inside one rectangular room. If the target and every walkable cell about to be read lie
  // in the interior of the same normal room - all floor by construction - the shortest 8-direction path between them
  // stays inside their bounding box, where nothing blocks a step (a diagonal's two orthogonal cells are in the box
  // too): the BFS label is the Chebyshev distance, and the map need not be extended at all (it stays suspended at
  // whatever level it has; labels never change once written). Not for a map that resumes on a walkability snapshot
  // of an earlier floor (cache_snap): its geometry is not this floor's.
  bool shortcut;
  {
    RG_PLANES(c);
    const int rid = room_of(c, tx, ty);
    bool ok = rid >= 0 && !((st->cache_snap >> slot) & 1u);
    RoomD rm = st->rooms[ok ? rid : 0];
    ok = ok && rm.kind == K_NORMAL && room_interior(rm, tx, ty);
    bool mine = true, walkable = false;
    if (ok && live && !oob) {
      walkable = can_walk(S[ny * c.W + nx]);
      mine = !walkable || room_interior(rm, nx, ny);
    }
    shortcut = ok && __all_sync(RG_FULL, mine);
    if (shortcut && walkable) nd = (uint32_t)max(abs(nx - tx), abs(ny - ty));
  }
  if (!shortcut) {
    bfs_extend(c, slot, mx, my);
    if (c.panic) return MV_CANT;
    if (live && !oob) nd = dm[ny * c.W + nx];
  }
  const bool reach = live && !oob && nd == 0 && can_move(c, mx, my, d, true);
  const bool cand = live && !oob && nd != 0xFFFFu && nd > 0;
  const uint32_t oobM = __ballot_sync(RG_FULL, oob);
  const uint32_t reachM = __ballot_sync(RG_FULL, reach);
  const int first_oob = oobM ? __ffs(oobM) - 1 : 99;
  const int first_reach = reachM ? __ffs(reachM) - 1 : 99;
  if (first_oob < first_reach) {  // `*dist_map.get_p(next)` out of range: the reference panics (:361)
    set_panic(c);
    return MV_CANT;
  }
  if (reachM) return MV_REACH;
  uint32_t key = cand ? ((nd << 4) | (uint32_t)d) : RG_FULL;  // stable sort_by_key + [0]
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(RG_FULL, key, o));
  if (key == RG_FULL) return MV_CANT;
  int bd = (int)(key & 15u);
  ox = mx + ddx(bd);
  oy = my + ddy(bd);
  return MV_CAN;
}
// rogue::Dungeon::move_enemy_randomly rogue/mod.rs:376-397
__device__ int move_enemy_randomly(Ctx& c, int mx, int my, uint32_t moved, int& ox, int& oy) {
  RG_PLANES(c);
  st->spec_req = 0;  // the dungeon stream moves on: a skeleton built from its old state is dead, ask again
  int d = (int)c.rd.range64(0, 8);
  int nx = mx + ddx(d), ny = my + ddy(d);
  if (cell_blocked(c, nx, ny, moved) || !can_move(c, mx, my, d, true)) return MV_CANT;
  if (nx == st->px && ny == st->py) return MV_REACH;
  ox = nx;
  oy = ny;
  return MV_CAN;
}

RG_DEV uint32_t clamp_parcent(int v) { return (uint32_t)min(100, max(0, v)); }  // Parcent::truncate rng.rs:152-154

// fight::roll fight.rs:52-72 on the enemy stream; returns -1 on a miss
__device__ int roll_dice(Rng& re, const int32_t* times, const int32_t* maxs, int n, uint32_t rate, int dam_plus) {
  bool did_hit = false;
  int sum = 0;
  for (int i = 0; i < n; ++i) {
    if (!re.parcent(rate)) continue;
    did_hit = true;
    int acc = 0;
    for (int t = 0; t < times[i]; ++t) acc += (int)re.range64(1, (uint64_t)(maxs[i] + 1));  // Dice::random, i64 lane
    sum += acc + dam_plus;
  }
  return did_hit ? sum : -1;
}

// actions::move_active_enemies + EnemyHandler::move_actives actions.rs:82-119, enemies.rs:366-424
// returns true when the player died (Some(Grave))
__device__ bool move_active_enemies(Ctx& c) {
  RG_PLANES(c);
  const rg_params& P = *c.P;
  uint32_t pending = 0;
  for (int m = 0; m < c.nrooms; ++m) {
    uint8_t f = st->mon[m].flags;
    if ((f & MF_PRESENT) && (f & MF_ACTIVE)) pending |= 1u << m;
  }
  if (!pending) return false;
  uint32_t moved = 0, attackers = 0;
  uint8_t attack_order[MAX_ROOMS];
  int n_att = 0;
  while (pending) {
    // next monster in BTreeMap order of its ORIGINAL position: [level, x, y] => x-major
    int best = -1;
    uint32_t bkey = RG_FULL;
    for (int m = 0; m < c.nrooms; ++m)
      if ((pending >> m) & 1u) {
        uint32_t key = ((uint32_t)st->mon[m].x << 8) | st->mon[m].y;
        if (key < bkey) { bkey = key; best = m; }
      }
    const int m = best;
    pending &= ~(1u << m);
    MonD mo = st->mon[m];
    const uint32_t attr = P.enemies[mo.kind].attr;
    bool randomly;
    if (c.re.does_happen(2) && (attr & EA_RANDOM)) randomly = true;
    else randomly = (!c.re.does_happen(5)) && (attr & EA_CONFUSED);
    int nx = mo.x, ny = mo.y, kind;
    if (randomly) kind = move_enemy_randomly(c, mo.x, mo.y, moved, nx, ny);
    else kind = move_enemy(c, mo.x, mo.y, st->px, st->py, moved, false, nx, ny);
    if (c.panic) return false;
    if (kind == MV_REACH) {
      attackers |= 1u << m;
      attack_order[n_att++] = (uint8_t)m;
    }
    if (kind != MV_CAN) {
      // re-insert at its own key: BTreeMap::insert replaces a monster that already moved there
      for (int o = 0; o < c.nrooms; ++o)
        if (((moved >> o) & 1u) && (st->mon[o].flags & MF_PRESENT) && st->mon[o].x == mo.x && st->mon[o].y == mo.y)
          st->mon[o].flags = 0;
    } else {
      st->mon[m].x = (uint8_t)nx;
      st->mon[m].y = (uint8_t)ny;
    }
    moved |= 1u << m;
  }
  if (!n_att) return false;
  st->quiet = 0;  // player.buttle()
  bool did_hit = false;
  for (int i = 0; i < n_att; ++i) {
    const MonD mo = st->mon[attack_order[i]];
    const rg_enemy_kind& k = P.enemies[mo.kind];
    int elevel = k.level + (int)lev_add(c);
    uint32_t rate = clamp_parcent((elevel + P.armor_def + 0 + 1) * 5);  // fight.rs:80-87, hit_prob_plus(10) = 0
    int dmg = roll_dice(c.re, k.dice_times, k.dice_max, (int)k.n_dice, rate, 0);  // damage_plus(10)+damage_plus(16) = 0
    if (dmg >= 0) {
      c.msg |= MSG_HIT_FROM;
      did_hit = true;
      st->hp = max(st->hp - dmg, 0);  // Player::get_damage player.rs:177-184
      if (st->hp == 0) {
        c.dead = 1;
        return true;
      }
    } else {
      c.msg |= MSG_MISS_FROM;
    }
  }
  if (did_hit) c.status_upd = 1;
  return false;
}

// Player::heal player.rs:221-240
__device__ bool heal(Ctx& c) {
  RG_PLANES(c);
  st->quiet += 1;
  int q = (int)st->quiet, lv = st->plevel, amount;
  if (lv < 8) amount = max(min(q + (lv << 1) - 20, 1), 0);
  else if (q >= 3) amount = (int)c.re.range64(1, (uint64_t)(lv - 6));
  else amount = 0;
  if (amount > 0) {
    st->hp = min(st->hp + amount, st->hp_max);
    st->quiet = 0;
    return true;
  }
  return false;
}
// actions::after_turn actions.rs:67-80 + Player::turn_passed player.rs:163-176.
// HOT: only the player's half runs here; the monster phase (move_active_enemies) is a separate
// kernel over the envs that have an active monster.
template <bool HOT>
__device__ bool after_turn(Ctx& c) {
  RG_PLANES(c);
  st->food_left -= 1;  // wraps like the release build
  if (st->food_left != 0) {
    uint32_t hunger = c.P->hunger_time / 10;
    bool hungry = st->food_left == hunger || st->food_left == hunger * 2;
    bool healed = heal(c);
    if (hungry || healed) c.status_upd = 1;
  }
  if constexpr (HOT) return false;
  else return move_active_enemies(c);
}

// Player::level_up + Leveling::check_level player.rs:185-197,346-352
__device__ bool level_up(Ctx& c, uint32_t gained) {
  RG_PLANES(c);
  const rg_params& P = *c.P;
  st->exp += gained;
  uint32_t cur = (uint32_t)(st->plevel - 1), diff = 0;
  if (cur < P.n_exps) {
    bool found = false;
    for (uint32_t i = cur; i < P.n_exps; ++i)
      if (st->exp < P.exps[i]) { diff = i - cur; found = true; break; }
    if (!found) { set_panic(c); return false; }
  }
  if (diff > 0) {
    st->plevel += (int)diff;
    int add = 0;
    for (uint32_t i = 0; i < diff; ++i) add += (int)c.re.range64(1, 11);
    st->hp_max += add;
    st->hp += add;
    return true;
  }
  return false;
}

// actions::move_player actions.rs:168-195 (incl. player_attack :140-166, get_item :206-231)
// returns `done`; *first_nonempty-style bookkeeping is left to the caller via flags
struct MoveOut {
  bool done;
  uint32_t redraw, status_upd, msg;
};
__device__ MoveOut move_player(Ctx& c, int d) {
  RG_PLANES(c);
  const rg_params& P = *c.P;
  MoveOut o{true, 0, 0, 0};
  const int px = st->px, py = st->py;
  // rogue::Dungeon::can_move_player
  if (!can_move(c, px, py, d, false)) return o;  // Notify(CantMove): not a message flag
  const int nx = px + ddx(d), ny = py + ddy(d);
  // EnemyHandler::get_cloned (enemies.rs:336-341): positions are unique between turns
  int target = -1;
  for (int m = 0; m < c.nrooms; ++m) {
    MonD mo = st->mon[m];
    if (target < 0 && (mo.flags & MF_PRESENT) && mo.x == nx && mo.y == ny) target = m;
  }
  if (target >= 0) {  // player_attack
    st->quiet = 0;
    MonD mo = st->mon[target];
    st->mon[target].flags = mo.flags | MF_ACTIVE;  // activate(): runs before the to-hit
    const rg_enemy_kind& k = P.enemies[mo.kind];
    int edef = k.defense - (int)lev_add(c);
    uint32_t rate = clamp_parcent((st->plevel + edef + 0 + 0 + P.weapon_hit_plus + 1) * 5);  // fight.rs:74-78
    int32_t t = P.weapon_times, mx = P.weapon_max;
    int dmg = roll_dice(c.re, &t, &mx, 1, rate, P.weapon_dam_plus);
    if (dmg >= 0) {
      o.msg |= MSG_HIT_TO;
      if (mo.hp <= dmg) {  // Enemy::get_damage enemies.rs:205-213
        st->mon[target].flags = 0;
        if (level_up(c, mo.exp)) o.status_upd = 1;
        o.msg |= MSG_KILLED;
        o.redraw = 1;
      } else {
        st->mon[target].hp = dmg - mo.hp;
      }
    } else {
      o.msg |= MSG_MISS_TO;
    }
    return o;
  }
  player_out(c, px, py);
  player_in(c, nx, ny, false);
  st->px = (int16_t)nx;
  st->py = (int16_t)ny;
  o.done = false;
  o.redraw = 1;
  const int pos = ny * c.W + nx;
  for (int i = 0; i < c.nrooms; ++i)
    if (st->item_pos[i] == pos && P.pack_accepts_gold) {
      st->gold += st->item_amt[i];
      st->item_pos[i] = 0xFFFF;
      o.status_upd = 1;  // Notify(GotItem) is not a message flag
      o.done = true;
    }
  return o;
}

// Floor::search floor.rs:349-370
__device__ void search(Ctx& c) {
  RG_PLANES(c);
  const int W = c.W;
  const int px = st->px, py = st->py;
  mark_rows(c, py - 1, py + 1);
  bool maps_done = false;
  for (int d = 0; d < 8; ++d) {
    int nx = px + ddx(d), ny = py + ddy(d);
    if (!inb(c, nx, ny)) continue;
    int idx = ny * W + nx;
    bool opened = false;
    if ((A[idx] & (A_HIDDEN | A_LOCKED)) && !maps_done) {  // walkability may change below
      snapshot_suspended_maps(c);
      maps_done = true;
      st->spec_req = 0;  // the dungeon stream is drawn from below (see move_enemy_randomly)
    }
    if ((A[idx] & A_HIDDEN) && c.rd.does_happen(c.P->passage_unlock_rate_inv)) {
      A[idx] = (A[idx] & (uint8_t)~(A_LOCKED | A_HIDDEN)) | A_VISIBLE;
      S[idx] = S_PASSAGE;
      opened = true;
    }
    if ((A[idx] & A_LOCKED) && c.rd.does_happen(c.P->door_unlock_rate_inv)) {
      A[idx] = (A[idx] & (uint8_t)~(A_LOCKED | A_HIDDEN)) | A_VISIBLE;
      S[idx] = S_DOOR;
      c.msg |= MSG_SECRET_DOOR;
      opened = true;
    }
    if (opened) {
      c.g_walk[ny * c.WW + (nx >> 5)] |= 1u << (nx & 31);
      c.s_dirty = 1;
    }
  }
  c.a_dirty = 1;
}

// RunTime::player_status core/src/lib.rs:345-356 -> Status::to_vec player.rs:417-430
__device__ void refresh_status(Ctx& c) {
  RG_PLANES(c);
  uint32_t hunger = c.P->hunger_time / 10;
  st->status[0] = (uint32_t)st->level;
  st->status[1] = st->gold;
  st->status[2] = (uint32_t)st->hp;
  st->status[3] = (uint32_t)st->hp_max;
  st->status[4] = 16;  // strength is hard-wired (player.rs:286)
  st->status[5] = 16;
  st->status[6] = 0;   // defense is never filled (player.rs:107-118)
  st->status[7] = (uint32_t)st->plevel;
  st->status[8] = st->exp;
  st->status[9] = st->food_left <= hunger ? 2u : (st->food_left <= hunger * 2 ? 1u : 0u);
}

// actions::process_action actions.rs:16-65. act: 0 Move, 1 MoveUntil, 2 Search, 3 DownStair.
// HOT: the caller has already routed "DownStair on a stair" to the generation kernel, so the
// floor generator is not compiled into the hot kernel at all.
template <bool HOT>
__device__ void process_action(Ctx& c, int act, int d) {
  RG_PLANES(c);
  bool ui = false;
  if (act == 3) {
    if (!HOT && S[st->py * c.W + st->px] == S_STAIR) {
      if constexpr (!HOT) gen_inl::new_level(&c, false);
      c.redraw = 1;
      c.status_upd = 1;
    } else {
      c.msg |= MSG_NO_DOWNSTAIR;
    }
    if (!c.panic) ui = after_turn<HOT>(c);
  } else if (act == 0) {
    MoveOut o = move_player(c, d);
    c.redraw |= o.redraw; c.status_upd |= o.status_upd; c.msg |= o.msg;
    if (!c.panic) ui = after_turn<HOT>(c);
  } else if (act == 1) {
    if constexpr (!HOT) {  // MoveUntil interleaves moves and monster phases: always the full-path kernel
      for (int guard = 0; guard < 512 && !c.panic; ++guard) {
        MoveOut o = move_player(c, d);
        int idx = st->py * c.W + st->px;
        // Dungeon::tile: the VISIBLE tile (rogue/mod.rs:321-328); keep running only on '.' / '#'
        bool on_open = (A[idx] & A_VISIBLE) && (S[idx] == S_FLOOR || S[idx] == S_PASSAGE);
        // the reactions of every sub-move matter only through these idempotent flags
        c.redraw |= o.redraw; c.status_upd |= o.status_upd; c.msg |= o.msg;
        if (o.done || !on_open) break;
        ui = after_turn<false>(c);
      }
    }
  } else if (act == 2) {
    search(c);
    c.redraw = 1;
    if (!c.panic) ui = after_turn<HOT>(c);
  }
  if (ui) st->ui_dead = 1;
}

// ------------------------------------------------------------------ observation compose
// RunTime::draw_screen core/src/lib.rs:264-285 + Dungeon::draw / draw_ranges / draw_enemy
// rogue/mod.rs:278-300,398-404 + Floor::history_map floor.rs:372-379.
// Whether monster `mo` is drawn for a player at (px, py): Dungeon::draw_enemy rogue/mod.rs:398-404
RG_DEV bool monster_shown(const Ctx& c, const MonD& mo, int px, int py) {
  const uint8_t* A = c.A;
  const EnvState* st = c.st;
  if (!(mo.flags & MF_PRESENT)) return false;
  const int idx = mo.y * c.W + mo.x;
  const bool vis = (A[idx] & (A_VISIBLE | A_DRAWN)) && mo.y >= 1 && mo.y < c.H - 1;
  const int dx = px - mo.x, dy = py - mo.y;
  bool show = dx * dx + dy * dy <= 2;  // Coord::is_adjacent
  if (!show) {                         // Floor::in_same_room floor.rs:381-393
    const int id = room_of(c, px, py);
    if (id >= 0 && room_of(c, mo.x, mo.y) == id) {
      const RoomD rm = st->rooms[id];
      show = (rm.kind == K_EMPTY) || (in_rect(rm, px, py) == in_rect(rm, mo.x, mo.y));
    }
  }
  return vis && show;
}

// The screen in HBM is persistent (the reference keeps PlayerState.map between redraws too), so a
// redraw recomputes only the rows that can differ from what is there: rows whose tile planes changed
// since the last compose (EnvState::dirty_rows, marked where the planes are written), rows where the
// last compose drew an overlay (ov_rows: monsters move without a redraw) and rows where one is drawn
// now. The result is what a full redraw gives; a typical move touches 3-6 of the 24 rows.
__device__ void compose(Ctx& c) {
  RG_PLANES(c);
  const int W = c.W, H = c.H;
  __syncwarp();
  const int px = st->px, py = st->py;
  // overlays first (positions only): monsters, items, player (core/src/lib.rs:272-283)
  int mon_idx = -1, item_idx = -1;
  uint64_t rows_now = 0;
  if (c.lane < c.nrooms) {
    const MonD mo = st->mon[c.lane];
    if (monster_shown(c, mo, px, py)) {
      mon_idx = mo.y * W + mo.x;
      rows_now |= 1ull << mo.y;
    }
    const uint16_t ip = st->item_pos[c.lane];
    if (ip != 0xFFFF && (A[ip] & (A_VISIBLE | A_DRAWN)) && ip >= W && ip < (H - 1) * W) {
      item_idx = ip;
      rows_now |= 1ull << (ip / W);
    }
  }
  const bool player_drawn = (A[py * W + px] & (A_VISIBLE | A_DRAWN)) && py >= 1 && py < H - 1;
  if (player_drawn) rows_now |= 1ull << py;
  {
    uint32_t lo = (uint32_t)rows_now, hi = (uint32_t)(rows_now >> 32);
    lo = __reduce_or_sync(RG_FULL, lo);
    hi = __reduce_or_sync(RG_FULL, hi);
    rows_now = ((uint64_t)hi << 32) | lo;
  }
  const uint64_t need = st->dirty_rows | st->ov_rows | rows_now;
  const int lo = W, hi = (H - 1) * W;  // rows 0 and H-1 are never written (python/src/lib.rs:44)
  // PARALLEL, 128-bit in / 128-bit out, 4 cells per op. When rows are whole 16-cell pieces (W % 16 == 0,
  // e.g. 80 = 5 pieces) the pieces of the needed rows are dealt to the lanes densely, so a typical
  // redraw (4-7 rows = 20-35 pieces) is one or two trips for the warp instead of four; otherwise every
  // lane walks its own pieces and skips the ones that lie in clean rows.
  const bool dense = (W & 15) == 0;
  const int cpr = W >> 4;
  const uint32_t need_lo = (uint32_t)need & (H >= 32 ? 0xFFFFFFFFu : ((1u << H) - 1u));
  const uint32_t need_hi = H > 32 ? (uint32_t)(need >> 32) & ((1u << (H - 32)) - 1u) : 0u;
  const int n_lo = __popc(need_lo);
  const int items = dense ? (n_lo + __popc(need_hi)) * cpr : c.CP / 16;
  for (int it = c.lane; it < items; it += 32) {
    int ch = it;
    if (dense) {
      const int ri = it / cpr;
      const int row = ri < n_lo ? (int)__fns(need_lo, 0, ri + 1) : 32 + (int)__fns(need_hi, 0, ri - n_lo + 1);
      ch = row * cpr + (it - ri * cpr);
    }
    const int base = ch * 16;
    if (!dense) {
      const int r0 = base / W, r1 = min(base + 15, c.C - 1) / W;  // a 16-cell piece lies in one or two rows
      if (!(((need >> r0) | (need >> r1)) & 1ull)) continue;
    }
    const uint4 s4 = *reinterpret_cast<const uint4*>(S + base);
    const uint4 a4 = *reinterpret_cast<const uint4*>(A + base);
    const uint32_t sv[4] = {s4.x, s4.y, s4.z, s4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
    uint32_t ov[4];
    uint32_t hbits = 0;
    const bool inside = base >= lo && base + 16 <= hi;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // Surface::tile (rogue/mod.rs:148-161) for four cells with one byte permute: the selector
      // nibbles are the surface codes, the 8-entry table is "#.-|%+^ "
      const uint32_t sel = (sv[q] & 7u) | ((sv[q] >> 4) & 0x70u) | ((sv[q] >> 8) & 0x700u) | ((sv[q] >> 12) & 0x7000u);
      const uint32_t tiles = __byte_perm(0x7C2D2E23u, 0x205E2B25u, sel);
      uint32_t vm = (av[q] >> 2) & 0x01010101u;  // IS_VISIBLE per byte (Cell::tile field.rs:92-98)
      if (!inside) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = base + q * 4 + k;
          if (idx < lo || idx >= hi) vm &= ~(1u << (8 * k));
        }
      }
      const uint32_t mask = (vm << 8) - vm;      // 0xFF in every visible byte
      ov[q] = (tiles & mask) | (0x20202020u & ~mask);
      hbits |= (((av[q] & 0x01010101u) * 0x01020408u) >> 24) << (4 * q);  // IS_VISITED bits (Floor::history_map)
    }
    *reinterpret_cast<uint4*>(c.g_screen + base) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    if (!c.hist_done) reinterpret_cast<uint16_t*>(c.g_hist)[ch] = (uint16_t)hbits;
  }
  __syncwarp();
  // overlays in increasing priority: a later store to the same cell wins
  if (mon_idx >= 0) c.g_screen[mon_idx] = (uint8_t)c.P->enemies[st->mon[c.lane].kind].tile;
  __syncwarp();
  if (item_idx >= 0) c.g_screen[item_idx] = '*';
  __syncwarp();
  if (c.lane == 0) {
    if (player_drawn) c.g_screen[py * W + px] = '@';
    // the step that leaves a floor shows the old floor's visited map (hist_done): the new floor's map is
    // written by the next compose, which therefore has to visit every row once more
    st->dirty_rows = c.hist_done ? ~0ull : 0ull;
    st->ov_rows = rows_now;
    if (c.g_rows) *c.g_rows |= need;
  }
  __syncwarp();
}

// GameConfig::build + GameStateImpl::reset + PlayerState::reset
// core/src/lib.rs:193-228, python/src/state_impls.rs:38-44, python/src/lib.rs:52-58
// CALLS selects the generator variant: true for the throughput kernels (k_reset, k_prefetch), false inside a step.
template <bool CALLS>
__device__ void reset_env(Ctx& c) {
  RG_PLANES(c);
  const rg_params& P = *c.P;
  c.rd.seed(st->seed);
  c.ri.seed(st->seed);
  c.re.seed(st->seed);
  if (!st->seeded) {  // seed: null => a fresh seed every reset (core/src/lib.rs:157-165)
    uint64_t z = ((uint64_t)st->seed[1] << 32 | st->seed[0]) + 0x9E3779B97F4A7C15ull * (uint64_t)(st->episode + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    uint64_t lo = P.has_seed_range ? P.seed_range_lo : 0, span = P.has_seed_range ? (P.seed_range_hi - P.seed_range_lo) : 0;
    if (span) z = lo + z % span;
    st->seed[0] = (uint32_t)z;
    st->seed[1] = (uint32_t)(z >> 32);
    st->seed[2] = st->seed[3] = 0;
  }
  st->episode += 1;
  st->level = 0;
  st->cache_n = 0;
  st->cache_head = 0;
  st->cache_snap = 0;
  st->hp = st->hp_max = P.init_hp;
  st->exp = 0;
  st->plevel = 1;
  st->food_left = P.hunger_time;
  st->quiet = 0;
  st->gold = P.init_gold;
  st->ui_dead = 0;
  st->error = 0;
  c.panic = 0;
  if constexpr (CALLS) gen_call::new_level(&c, true);
  else gen_inl::new_level(&c, true);
  refresh_status(c);
  c.hist_done = 0;
  st->message = 0;
  st->is_terminal = 0;
  st->steps = 0;
}

// KeyMap::ai input.rs:74-99 ; act 4 = NoOp, -1 = not in the map
RG_DEV int map_key(uint8_t key, int& d) {
  d = D_STAY;
  switch (key) {
    case 'l': d = D_RIGHT; return 0;
    case 'k': d = D_UP; return 0;
    case 'j': d = D_DOWN; return 0;
    case 'h': d = D_LEFT; return 0;
    case 'u': d = D_RIGHTUP; return 0;
    case 'y': d = D_LEFTUP; return 0;
    case 'n': d = D_RIGHTDOWN; return 0;
    case 'b': d = D_LEFTDOWN; return 0;
    case 'L': d = D_RIGHT; return 1;
    case 'K': d = D_UP; return 1;
    case 'J': d = D_DOWN; return 1;
    case 'H': d = D_LEFT; return 1;
    case 'U': d = D_RIGHTUP; return 1;
    case 'Y': d = D_LEFTUP; return 1;
    case 'N': d = D_RIGHTDOWN; return 1;
    case 'B': d = D_LEFTDOWN; return 1;
    case 's': return 2;
    case '>': return 3;
    case '.': return 4;
    default: return -1;
  }
}

}  // namespace rg
