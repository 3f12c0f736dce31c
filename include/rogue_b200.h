/* include/rogue_b200.h — C ABI of the B200 batched Rogue-Gym simulator.
 *
 * This is the drop-in boundary for the reference's extension module
 * `rogue_gym_python._rogue_gym` (reference: python/src/lib.rs, state_impls.rs,
 * thread_impls.rs). Everything behind it is hand-written CUDA for sm_100a; there is no
 * CPU fallback: every entry point that computes fails with RG_ERR_CUDA when no device or
 * kernel image is available.
 *
 * Conventions: plain pointers and sizes only; 0 = success, nonzero = rg_status with a
 * message in rg_last_error(); one caller thread per handle; calls are asynchronous on the
 * batch's own CUDA stream unless they return host data.
 *
 * Reference interface each entry point replaces (paths relative to /root/reference):
 *   rg_parse_config     GameConfig::from_json                 core/src/lib.rs:144-146
 *                       + GameConfig::symbol_max              core/src/lib.rs:150-155
 *                       + size checks of to_global            core/src/lib.rs:166-184
 *   rg_create           ParallelGameState::new / GameState::__new__ (build every env)
 *                                                             python/src/lib.rs:214-224,267-296
 *                       = ThreadConductor::new                python/src/thread_impls.rs:14-35
 *   rg_seed             ParallelGameState::seed / GameState::set_seed
 *                                                             python/src/lib.rs:229-232,303-309
 *                       = Instruction::Seed                   python/src/thread_impls.rs:125-128
 *   rg_reset            ParallelGameState::reset / GameState::reset -> GameStateImpl::reset
 *                                                             python/src/state_impls.rs:38-44
 *   rg_step             ParallelGameState::step -> ThreadConductor::step (auto_reset = 1)
 *                                                             python/src/thread_impls.rs:61-81
 *                       GameState::react -> GameStateImpl::react (auto_reset = 0)
 *                                                             python/src/state_impls.rs:51-79
 *   rg_step_host        the same call with host buffers (what the PyO3 layer does per call:
 *                       Vec<u8> in, Vec<PlayerState> out)     python/src/lib.rs:315-321
 *   rg_views / rg_fetch ParallelGameState::states / GameState::prev
 *                                                             python/src/lib.rs:238-241,310-314
 *   rg_step_mirror      the same host-facing step; the Vec<PlayerState> it replaces is kept current
 *                       in a host mirror by delta writes instead of a full copy (no reference
 *                       counterpart for the mechanism)          python/src/lib.rs:315-321
 *   rg_encode           PlayerState::{gray_image,symbol_image}[_with_hist]
 *                                                             python/src/lib.rs:158-205
 *   rg_status_vec order StatusFlagInner::to_vector            python/src/flags.rs:63-85
 *   rg_dump             (no reference counterpart: parity harness)
 */
#ifndef ROGUE_B200_H
#define ROGUE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_MAX_ENEMY_KINDS 32
#define RG_MAX_DICE 4
#define RG_MAX_EXPS 32
#define RG_MAX_INIT_DRAWS 8
#define RG_MAX_ROOMS 16
#define RG_DIST_CACHE 9

typedef enum rg_status {
  RG_OK = 0,
  RG_ERR_INVALID_INPUT = 1, /* ErrorKind::InvalidInput: key outside KeyMap::ai (input.rs:74-99) */
  RG_ERR_IGNORED_INPUT = 2, /* ErrorKind::IgnoredInput: action after death (core/src/lib.rs:314) */
  RG_ERR_PANIC = 3,         /* a state the reference panics in; sticky per env */
  RG_ERR_SETTING = 4,       /* ErrorKind::InvalidSetting / unsupported configuration */
  RG_ERR_PARSE = 5,         /* "Failed to parse config" (python/src/lib.rs:219,274-277) */
  RG_ERR_CUDA = 6,          /* no device, no kernel image, or a CUDA runtime failure */
  RG_ERR_ARG = 7
} rg_status;

/* One monster kind in EnemyHandler order (stable sort by rarelity, enemies.rs:251). */
typedef struct rg_enemy_kind {
  int32_t tile;
  int32_t level;
  int32_t defense;
  uint32_t exp;
  uint32_t attr; /* EnemyAttr bits, enemies.rs:125-137 */
  uint32_t n_dice;
  int32_t dice_times[RG_MAX_DICE];
  int32_t dice_max[RG_MAX_DICE];
} rg_enemy_kind;

/* Flat form of GameConfig: exactly what the device path reads. */
typedef struct rg_params {
  int32_t width, height;
  int32_t room_num_x, room_num_y, min_room_x, min_room_y;
  uint32_t max_empty_rooms, amulet_level, maze_rate_inv, dark_level;
  uint32_t hidden_passage_rate_inv, locked_door_rate_inv, max_extra_edges;
  uint32_t door_unlock_rate_inv, passage_unlock_rate_inv;
  uint32_t gold_rate_inv, gold_base, gold_per_level, gold_minimum;
  uint32_t hunger_time;
  int32_t init_hp;
  uint32_t n_exps;
  uint32_t exps[RG_MAX_EXPS];
  int32_t pack_accepts_gold;
  uint32_t init_gold;
  int32_t weapon_times, weapon_max, weapon_hit_plus, weapon_dam_plus;
  int32_t armor_def;
  uint32_t n_init_draws;
  uint32_t init_draw_lo[RG_MAX_INIT_DRAWS], init_draw_hi[RG_MAX_INIT_DRAWS];
  uint32_t n_enemies;
  rg_enemy_kind enemies[RG_MAX_ENEMY_KINDS];
  uint32_t appear_rate_gold, appear_rate_nogold;
  int32_t hide_dungeon;
  uint32_t symbols;
  /* seed handling (core/src/lib.rs:157-165) */
  int32_t has_seed;
  uint64_t seed_lo, seed_hi;
  int32_t has_seed_range;
  uint64_t seed_range_lo, seed_range_hi; /* low 64 bits of the half-open range */
} rg_params;

typedef struct rg_batch rg_batch;

/* Device-resident observation block; pointers stay valid and fixed for the batch's life. */
typedef struct rg_views {
  int64_t n_envs;
  int32_t width, height;
  int32_t cell_stride;   /* bytes between envs in screen (W*H rounded up to 16) */
  int32_t hist_stride;   /* bytes between envs in history_bits */
  uint8_t* screen;       /* [N][cell_stride] ASCII tiles, row-major H x W (PlayerState.map) */
  uint8_t* history_bits; /* [N][hist_stride] bit y*W+x = visited (PlayerState.history) */
  uint32_t* status;      /* [N][10] Status::to_vec order (player.rs:417-430) */
  int32_t* reward;       /* [N] max(0, gold - gold_before) as ParallelRogueEnv.step (parallel.py:60-63) */
  uint8_t* done;         /* [N] PlayerState.is_terminal */
  uint32_t* message;     /* [N] MessageFlagInner bits (python/src/flags.rs:9-17) */
  uint8_t* error;        /* [N] rg_status of the last step / sticky panic */
} rg_views;

/* Host mirror filled by rg_step_host / rg_fetch. Any pointer may be NULL to skip it. */
typedef struct rg_host_obs {
  uint8_t* screen;       /* [N][W*H] (dense, no padding) */
  uint8_t* history;      /* [N][W*H] 0/1 bytes */
  uint32_t* status;      /* [N][10] */
  int32_t* reward;       /* [N] */
  uint8_t* done;         /* [N] */
  uint32_t* message;     /* [N] */
  uint8_t* error;        /* [N] */
} rg_host_obs;

/* Canonical full state of one env, same layout the oracle exposes, for parity tests. */
typedef struct rg_dump_scalars {
  int32_t level;
  int32_t px, py;
  int32_t hp, hp_max;
  uint32_t exp;
  int32_t plevel;
  uint32_t food_left, quiet;
  uint32_t gold;
  int32_t ui_dead;
  int32_t steps;
  int32_t is_terminal;
  uint32_t message;
  int32_t error;
  int32_t n_monsters, n_items, n_cache;
  uint32_t status[10];
  uint32_t rng[12];
} rg_dump_scalars;

typedef struct rg_dump {
  rg_dump_scalars s;
  uint8_t* surface;  /* [W*H] Surface codes (rogue/mod.rs:137-146 declaration order) */
  uint8_t* attr;     /* [W*H] CellAttr bits 0..5, bit 6 = member of Floor.doors */
  int32_t* monsters; /* [RG_MAX_ROOMS][8] x,y,kind,hp,active,level,defense,exp in (x,y) order */
  int32_t* items;    /* [RG_MAX_ROOMS][3] x,y,amount in y*W+x order */
  int32_t* cache_xy; /* [RG_DIST_CACHE][2] FIFO order, -1 padded */
  uint16_t* cache_maps; /* [RG_DIST_CACHE][W*H] or NULL */
  int32_t* rooms;    /* [RG_MAX_ROOMS][8] kind,is_dark,is_visited,has_gold,x0,y0,x1,y1 */
} rg_dump;

/* ---- configuration (host only, no device needed) */
int rg_parse_config(const char* json, rg_params* out, char* err, size_t err_len);
int rg_validate_params(const rg_params* p, char* err, size_t err_len);

/* ---- lifetime */
/* n_cfg == 1 broadcasts cfg_json[0] to all envs; n_cfg == n_envs gives every env its own JSON
 * (python/src/lib.rs:270-280). The configs of one batch must agree on width, height, room_num_x,
 * room_num_y and the symbol count (memory layout, observation shape) and must all set or all omit
 * "seed"; everything else - monsters, rates, gold, player, hide_dungeon, seed - may differ per env. */
int rg_create(const char* const* cfg_json, int64_t n_cfg, int64_t n_envs, int64_t max_steps, int device,
              rg_batch** out);
int rg_create_from_params(const rg_params* p, int64_t n_envs, int64_t max_steps, int device, rg_batch** out);
void rg_destroy(rg_batch* b);
const char* rg_last_error(rg_batch* b); /* b may be NULL: last create/parse error of this thread */
const char* rg_version(void);

/* ---- stepping */
int rg_seed(rg_batch* b, const uint64_t* seed_lo, const uint64_t* seed_hi /* nullable */);
/* Seeds only the first `count` envs (arrays of `count` entries); the others keep the seed they have:
 * ThreadConductor::seed zips workers with seeds (python/src/thread_impls.rs:45-50). */
int rg_seed_first(rg_batch* b, const uint64_t* seed_lo, const uint64_t* seed_hi /* nullable */, int64_t count);
int rg_reset(rg_batch* b);
int rg_step(rg_batch* b, const uint8_t* actions_dev, int auto_reset);
int rg_step_host(rg_batch* b, const uint8_t* actions_host, int auto_reset, rg_host_obs* out);
/* Makes the batch's stream wait for the background generation of next-episode games queued so
 * far (asynchronous; used by benchmarks so that a timed region ends with no work in flight). */
int rg_quiesce(rg_batch* b);
/* Event counters since creation: [0] episode ends served by a prefetched game, [1] episode ends
 * generated synchronously, [2] steps taken on the full path (descents, MoveUntil), [3] games built in
 * the background, [4] prefetched games found stale, [5] env-steps with an active monster, [6] spare,
 * [7] env-steps finished by the thread-per-env kernel (k_step_fast). */
int rg_stats(rg_batch* b, uint64_t* out8);
/* Timeline of the last 512 steps (only when the batch was created with RG_TRACE=1 in the environment):
 * out[512][12][2] = first start / last end (globaltimer ns) of kernel k of step slot s; k = 0 player kernel of the
 * active-monster envs, 1 their monster kernel, 2 thread-per-env kernel, 3 full path, 4 synchronous resets, 5 background
 * prefetch, 6 / 7 player / monster kernel of the thread kernel's leftovers, 8 / 9 host-mirror passes, 10 scan.
 * Reading clears the slots. */
int rg_trace(rg_batch* b, uint64_t* out, int64_t* steps_launched);
int rg_sync(rg_batch* b);          /* waits for the stream and raises per-env errors like the reference */
int rg_views_get(rg_batch* b, rg_views* out);
int rg_fetch(rg_batch* b, rg_host_obs* out);
/* The states' own terminal flags, out_host u8 [N] (ParallelGameState::states -> Instruction::State,
 * python/src/thread_impls.rs:51-60,131): after an auto-reset step `done` is 1 - the conductor flags the copy it
 * returns (thread_impls.rs:69-79) - while the env's state, the fresh game, is not terminal. */
int rg_fetch_terminal(rg_batch* b, uint8_t* out_host);
/* What happens to an env that reaches a state in which the reference panics (the reference loses
 * the worker thread and with it the whole conductor, python/src/thread_impls.rs:111-135):
 *   0 sticky (default)  the env is frozen; every later step reports RG_ERR_PANIC for it
 *   1 terminal          on auto-reset steps the step that hits the state reports done = 1 together with
 *                       RG_ERR_PANIC once, and the env goes on with a fresh episode (its next seed)
 * Also settable at creation through the environment variable RG_PANIC_POLICY=sticky|terminal. */
int rg_set_panic_policy(rg_batch* b, int policy);
/* ---- trainer-facing step: everything a policy loop needs from one call, all on the device and on the
 * batch's stream (what ParallelRogueEnv.step + StairRewardParallel + ImageSetting.expand per state do in
 * Python: python/rogue_gym/envs/parallel.py:44-66, wrappers.py:44-64, rogue_env.py:84-98).
 * action_idx_dev: [N] indices into the gym action table ". h j k l n b u y > s", index_bytes = 1, 4 or 8
 * (uint8 / int32 / int64), or index_bytes = 0 for uint8 ASCII keys as rg_step takes them (capitals = move
 * until blocked); an index outside 0..10 gives that env RG_ERR_INVALID_INPUT. obs_out_dev: f32
 * [N][channels][H][W] image of the new state (mode / status_flag / with_hist as rg_encode; NULL = no image);
 * reward_out_dev: f32 [N] = gold gained, plus stair_reward on the step an env gets deeper than the level
 * last seen for it (NULL = skip). Steps with auto-reset; done / status / screen are in rg_views. */
int rg_step_train(rg_batch* b, const void* action_idx_dev, int index_bytes, int mode, uint32_t status_flag, int with_hist,
                  float* obs_out_dev, float* reward_out_dev, float stair_reward);
/* Forget the levels rg_step_train has seen (call after rg_reset). */
int rg_train_reset(rg_batch* b);

/* ---- host mirror: the observation block kept current in HOST memory without copying it whole.
 * rg_mirror_get allocates (once) pinned host buffers that the device can write, owned by the batch and
 * valid until rg_destroy: out->screen [N][W*H], status, reward, done, message, error as in rg_host_obs;
 * out->history is NULL - the visited map is mirrored bit-packed, *history_bits [N][hist_stride] with
 * bit y*W+x of an env's row = visited (the layout of rg_views.history_bits), and only from the first call that
 * passes a non-NULL history_bits on (it costs about a third of a step's PCIe writes).
 * rg_mirror_sync brings the mirror up to date with the device block: kernels compare the block with a
 * device-side shadow of what the host holds and store only the 64-byte lines that changed, over PCIe,
 * into the mirror; it returns once the last of those writes has landed, so the host may read at once.
 * rg_step_mirror = actions H2D + rg_step + rg_mirror_sync: the host-facing step of rg_step_host
 * (python/src/lib.rs:315-321) at a fraction of its PCIe traffic. *bytes_to_host (nullable) receives the
 * bytes the call stored into the mirror. The caller must not write to the mirror. */
int rg_mirror_get(rg_batch* b, rg_host_obs* out, uint8_t** history_bits);
int rg_mirror_sync(rg_batch* b, uint64_t* bytes_to_host);
int rg_step_mirror(rg_batch* b, const uint8_t* actions_host, int auto_reset, uint64_t* bytes_to_host);
void* rg_stream(rg_batch* b);      /* cudaStream_t the batch launches on */
int64_t rg_launch_count(rg_batch* b); /* kernels launched so far by this batch */

/* ---- observation encoders: out_dev is f32 [N][channels][H][W]; mode 0 gray, 1 symbol.
 * Returns channels through *channels. InvalidTileError (symbol.rs:62-64) sets error[env]. */
int rg_encode(rg_batch* b, int mode, uint32_t status_flag, int with_hist, float* out_dev, int* channels);
int rg_encode_channels(const rg_batch* b, int mode, uint32_t status_flag, int with_hist);
/* Compact observation: what symbol_image carries, without the one-hot expansion (1.9 KB instead of 401 KB per env
 * for the reference-default ImageSetting; a policy embeds / one-hots the ids on the device, see
 * rogue_gym.envs.device.SymbolExpand). sym_out_dev u8 [N][W*H]: Symbol::from_tile of every screen cell
 * (symbol.rs:17-40), 0 .. symbols-1; a tile without a symbol sets RG_ERR_SETTING in error[env] and is written as 0.
 * (The id symbols-1 - the largest monster tile - is passed through: gray_image encodes it, symbol_image raises
 * InvalidTileError for it, symbol.rs:60-64; SymbolExpand leaves its plane empty.) status_out_dev i32 [N][9]: the status vector in StatusFlagInner order (flags.rs:63-85), nullable.
 * hist_out_dev u8 [N][W*H] 0/1 visited map, nullable. Device buffers, on the batch's stream. */
int rg_encode_compact(rg_batch* b, uint8_t* sym_out_dev, int32_t* status_out_dev, uint8_t* hist_out_dev);
/* The same encoders for n detached PlayerState values held on the host (what PlayerState.gray_image /
 * symbol_image do per object, python/src/lib.rs:158-205): screens [n][W*H], history [n][W*H] 0/1 bytes
 * (may be NULL unless with_hist), status [n][10]; out_host f32 [n][channels][H][W]. Runs on the device. */
int rg_encode_states(rg_batch* b, int64_t n, const uint8_t* screens, const uint8_t* history, const uint32_t* status,
                     int mode, uint32_t status_flag, int with_hist, float* out_host, int* channels);

/* ---- parity harness */
int rg_dump_env(rg_batch* b, int64_t env, rg_dump* out);
int rg_state_hash(rg_batch* b, uint64_t* out_host /* [N] */);
/* Dungeon::move_enemy with an always-false skip, for the known-answer test
 * (core/src/dungeon/rogue/mod.rs:566-578): 0 CantMove, 1 CanMove, 2 Reach. */
int rg_test_move_enemy(rg_batch* b, int64_t env, int fx, int fy, int tx, int ty, int* kind, int* nx, int* ny);
/* The tile planes and room tables of every env, into caller-owned DEVICE buffers (any may be NULL), on the
 * batch's stream: surface_dev / attr_dev u8 [N][W*H] dense (codes as in rg_dump), rooms_dev i16
 * [N][RG_MAX_ROOMS][8] = kind (0 normal, 1 maze, 2 empty), is_dark, x0, y0, x1, y1 (half-open, walls
 * included), player x, player y. For the reference's generator property tests restated on whole batches
 * (passages.rs:343-379 connectivity, rooms.rs:308-340 pos_check, floor.rs:466-488 secret_door). */
int rg_export_floors(rg_batch* b, uint8_t* surface_dev, uint8_t* attr_dev, int16_t* rooms_dev);

#ifdef __cplusplus
}
#endif
#endif
