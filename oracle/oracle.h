/* oracle/oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. This is a serial CPU restatement of the reference's
 * (kngwyu/rogue-gym @ c78608b) step / reset / observe path. It exists to check the
 * CUDA product, and to serve as the timed CPU baseline in bench.py. Nothing in the
 * product package may include, link or call it.
 *
 * Parity status: PINNED for dungeon generation, gold placement, player movement,
 * MoveUntil, stairs, status and rewards (reference goldens: python/tests/data.py:83-108,
 * python/tests/test_ff_env.py:14-22, python/tests/test_st_env.py:27-37,
 * core/src/dungeon/rogue/mod.rs:566-578; see tests/test_oracle_golden.py).
 * UNPINNED (no reference output exists; source-as-read): monster spawn / AI / combat,
 * heal, hunger, search, seed 0, the x=0 maze panic.
 */
#ifndef ROGUE_ORACLE_H
#define ROGUE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_ENEMY_KINDS 32
#define ORC_MAX_DICE 4
#define ORC_MAX_EXPS 32
#define ORC_MAX_INIT_DRAWS 8
#define ORC_MAX_ROOMS 16
#define ORC_DIST_CACHE 9

/* One monster kind, already in EnemyHandler order (stable sort by rarelity,
 * core/src/character/enemies.rs:251). */
typedef struct orc_enemy_kind {
  int32_t tile;
  int32_t level;
  int32_t defense;
  uint32_t exp;
  uint32_t attr;     /* EnemyAttr bits, enemies.rs:125-137 */
  uint32_t n_dice;
  int32_t dice_times[ORC_MAX_DICE];
  int32_t dice_max[ORC_MAX_DICE];
} orc_enemy_kind;

/* Flat form of GameConfig (core/src/lib.rs:42-86) restricted to what the path reads. */
typedef struct orc_params {
  int32_t width, height;
  int32_t room_num_x, room_num_y, min_room_x, min_room_y;
  uint32_t max_empty_rooms, amulet_level, maze_rate_inv, dark_level;
  uint32_t hidden_passage_rate_inv, locked_door_rate_inv, max_extra_edges;
  uint32_t door_unlock_rate_inv, passage_unlock_rate_inv;
  uint32_t gold_rate_inv, gold_base, gold_per_level, gold_minimum;
  uint32_t hunger_time;
  int32_t init_hp;
  uint32_t n_exps;
  uint32_t exps[ORC_MAX_EXPS];
  int32_t pack_accepts_gold; /* a Gold stack or a free slot exists (itembox.rs:30-40) */
  uint32_t init_gold;
  int32_t weapon_times, weapon_max, weapon_hit_plus, weapon_dam_plus; /* wielded weapon */
  int32_t armor_def;                                                  /* Player::arm() */
  uint32_t n_init_draws;                                              /* weapon.rs:159 */
  uint32_t init_draw_lo[ORC_MAX_INIT_DRAWS], init_draw_hi[ORC_MAX_INIT_DRAWS];
  uint32_t n_enemies;
  orc_enemy_kind enemies[ORC_MAX_ENEMY_KINDS];
  uint32_t appear_rate_gold, appear_rate_nogold;
  int32_t hide_dungeon;
  uint32_t symbols;
} orc_params;

/* Full internal state, canonical layout shared with the product's rg_dump so that
 * tests can compare field by field. */
typedef struct orc_scalars {
  int32_t level;
  int32_t px, py;
  int32_t hp, hp_max;
  uint32_t exp;
  int32_t plevel;
  uint32_t food_left, quiet;
  uint32_t gold;       /* pack gold */
  int32_t ui_dead;     /* runtime UI is the grave modal */
  int32_t steps;
  int32_t is_terminal;
  uint32_t message;
  int32_t error;       /* sticky error: 0 none, else ORC_ERR_* */
  int32_t n_monsters, n_items, n_cache;
  uint32_t status[10]; /* displayed status, Status::to_vec order */
  uint32_t rng[12];    /* dungeon, item, enemy × (x,y,z,w) */
} orc_scalars;

enum {
  ORC_OK = 0,
  ORC_ERR_INVALID_INPUT = 1, /* ErrorKind::InvalidInput */
  ORC_ERR_IGNORED_INPUT = 2, /* ErrorKind::IgnoredInput (input after death) */
  ORC_ERR_PANIC = 3,         /* a Rust panic in the reference (worker thread dies) */
  ORC_ERR_SETTING = 4        /* invalid setting / generation failure */
};

void* orc_create(const orc_params* p, int64_t max_steps);
void orc_destroy(void* env);
void orc_set_seed(void* env, uint64_t lo, uint64_t hi);
/* GameStateImpl::reset (python/src/state_impls.rs:38-44) */
int orc_reset(void* env);
/* GameStateImpl::react (python/src/state_impls.rs:51-79) */
int orc_react(void* env, uint8_t key);
/* ThreadConductor::step for one worker (python/src/thread_impls.rs:61-81): react,
 * then auto-reset when terminal and keep is_terminal = 1 on the fresh state. */
int orc_step_auto(void* env, uint8_t key);
const char* orc_last_error(void* env);

/* observation = PlayerState (python/src/lib.rs:31-38) */
void orc_get_obs(void* env, uint8_t* screen /*H*W*/, uint8_t* history /*H*W*/, uint32_t* status10,
                 uint32_t* message, int32_t* is_terminal);
/* internals */
void orc_get_scalars(void* env, orc_scalars* out);
void orc_get_grid(void* env, uint8_t* surface /*H*W*/, uint8_t* attr /*H*W, bit6 = in doors set*/);
/* monsters in (x,y) order: x,y,kind,hp,active,level,defense,exp ; items in y*W+x order: x,y,amount */
void orc_get_entities(void* env, int32_t* monsters /*[ORC_MAX_ROOMS][8]*/, int32_t* items /*[ORC_MAX_ROOMS][3]*/);
/* DistCache: coords in FIFO order and, if maps != NULL, the u16 maps (0xFFFF = unreachable) */
void orc_get_dist_cache(void* env, int32_t* xy /*[9][2]*/, uint16_t* maps /*[9][H*W] or NULL*/);
/* rooms: kind(0 normal,1 maze,2 empty), is_dark, is_visited, has_gold, x0,y0,x1,y1 */
void orc_get_rooms(void* env, int32_t* rooms /*[ORC_MAX_ROOMS][8]*/);
/* counters for measurement: raw next_u32 calls per stream since reset */
void orc_get_draw_counts(void* env, uint64_t* counts3);

/* f32 encoders (python/src/lib.rs:72-111,158-205). mode 0 = gray, 1 = symbol.
 * Returns number of channels written, or -1 on InvalidTileError. */
int orc_encode(void* env, int mode, uint32_t status_flag, int with_hist, float* out);

/* Known-answer hook: Dungeon::move_enemy with an always-false skip
 * (core/src/dungeon/rogue/mod.rs:566-578). Returns 0 CantMove, 1 CanMove (nx,ny set), 2 Reach. */
int orc_test_move_enemy(void* env, int fx, int fy, int tx, int ty, int* nx, int* ny);

/* Batch helper used as the timed CPU baseline: steps n envs for `steps` turns with the
 * synthetic action stream of SURVEY.md §8d on `threads` host threads (static partition).
 * Returns elapsed seconds; digest (xor of per-env state hashes) is written to *digest. */
double orc_batch_rollout(void** envs, int64_t n, int64_t first_env_id, int64_t t0, int64_t steps,
                         int threads, int compose_only_on_redraw, uint64_t* digest);
uint64_t orc_state_hash(void* env);
/* hooks for the reference's own unit tests (fenwick.rs:262-305, passages.rs:272-296) */
int64_t orc_test_ordset(uint64_t cap, const uint64_t* members, const uint8_t* ops, uint8_t* ops_result, int64_t n,
                        int64_t k, uint64_t* len_out);
int orc_test_ordset_from_range_contains(uint64_t lo, uint64_t hi, uint64_t e);
int orc_test_edges(int x0, int x1, int y0, int y1, int direction, int inclusive, int32_t* out_xy, int cap);

/* Lock-step helpers for the parity tests: one call steps / resets / reads n envs. */
void orc_batch_step(void** envs, int64_t n, const uint8_t* keys, int auto_reset, int threads, int32_t* rc_out);
void orc_batch_reset(void** envs, int64_t n, int threads, int32_t* rc_out);
void orc_batch_get_obs(void** envs, int64_t n, uint8_t* screen, uint8_t* history, uint32_t* status, uint32_t* message,
                       uint8_t* is_terminal);
void orc_batch_hash(void** envs, int64_t n, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif
