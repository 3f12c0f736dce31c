// rg_config.cpp — GameConfig JSON -> flat rg_params (host only).
//
// Replaces GameConfig::from_json (core/src/lib.rs:144-146) and the parts of
// GameConfig::build that only depend on the configuration: the monster table in
// EnemyHandler order (enemies.rs:245-262), the player's wielded weapon / worn armour
// (player.rs:198-219), the item-stream draws of Player::init_items (item/mod.rs:412-422,
// weapon.rs:159), symbol_max (core/src/lib.rs:150-155) and the size checks of to_global
// (core/src/lib.rs:166-184). Schema: core/src/lib.rs:42-86, dungeon/rogue/mod.rs:22-61,
// character/player.rs:16-32, character/enemies.rs:17-27,110-120, item/gold.rs:7-16,
// item/weapon.rs:130-141, item/armor.rs:128-140, item/mod.rs:160-180.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rogue_b200.h"
#include "rg_json.h"

namespace {

using rgjson::Value;

struct KindRow {
  int tile;
  std::vector<std::pair<int, int>> dice;
  uint32_t attr;
  int defense;
  uint32_t exp;
  int level;
  int rarelity;
};

enum : uint32_t {
  MEAN = 1u << 0, FLYING = 1u << 1, REGENERATE = 1u << 2, GREEDY = 1u << 3, INVISIBLE = 1u << 4,
  RUSTS_ARMOR = 1u << 5, STEAL_GOLD = 1u << 6, REDUCE_STR = 1u << 7, FREEZES = 1u << 8, RANDOM = 1u << 9
};

// BUILTIN_ENEMIES (enemies.rs:474-761), index = preset id
const KindRow& builtin_enemy(size_t i) {
  static const KindRow T[26] = {
      {'A', {{0, 0}}, MEAN | RUSTS_ARMOR, 2 | 8, 20, 5, 12},
      {'B', {{1, 2}}, FLYING | RANDOM, 3, 1, 1, 2},
      {'C', {{1, 2}, {1, 5}, {1, 5}}, 0, 4, 17, 4, 10},
      {'D', {{1, 8}, {1, 8}, {3, 10}}, MEAN, 3, 5000, 10, 25},
      {'E', {{1, 2}}, MEAN, 7, 2, 1, 1},
      {'F', {}, MEAN, 3, 80, 8, 15},
      {'G', {{4, 3}, {3, 5}}, FLYING | MEAN | REGENERATE, 2, 2000, 13, 23},
      {'H', {{1, 8}}, MEAN, 5, 3, 1, 4},
      {'I', {{0, 0}}, FREEZES, 9, 5, 1, 5},
      {'J', {{2, 12}, {2, 4}}, 0, 6, 3000, 15, 24},
      {'K', {{1, 4}}, MEAN, 7, 1, 1, 0},
      {'L', {{1, 1}}, STEAL_GOLD, 8, 10, 3, 9},
      {'M', {{3, 4}, {3, 4}, {2, 5}}, MEAN, 2, 200, 8, 21},
      {'N', {{0, 0}}, 0, 9, 37, 3, 13},
      {'O', {{1, 8}}, GREEDY, 6, 5, 1, 7},
      {'P', {{4, 4}}, INVISIBLE, 3, 120, 8, 18},
      {'Q', {{1, 5}, {1, 5}}, MEAN, 3, 15, 3, 11},
      {'R', {{1, 6}}, REDUCE_STR | MEAN, 3, 9, 2, 6},
      {'S', {{1, 3}}, MEAN, 5, 2, 1, 3},
      {'T', {{1, 8}, {1, 8}, {2, 6}}, MEAN | REGENERATE, 4, 120, 6, 16},
      {'U', {{1, 9}, {1, 9}, {2, 9}}, MEAN, -2, 190, 7, 20},
      {'V', {{1, 19}}, MEAN | REGENERATE, 1, 350, 8, 22},
      {'W', {{1, 6}}, 0, 4, 55, 5, 17},
      {'X', {{4, 4}}, 0, 7, 100, 7, 19},
      {'Y', {{1, 6}, {1, 6}}, 0, 6, 50, 4, 14},
      {'Z', {{1, 8}}, MEAN, 8, 6, 2, 8},
  };
  return T[i];
}

struct WeaponRow {
  std::string name;
  int times, max;
  uint32_t num_lo, num_hi;
};
// BUILTIN_WEAPONS (weapon.rs:196-296)
const std::vector<WeaponRow>& builtin_weapons() {
  static const std::vector<WeaponRow> T = {
      {"mace", 2, 4, 1, 2},  {"long-sword", 3, 4, 1, 2},       {"bow", 1, 1, 1, 2},   {"arrow", 1, 1, 8, 17},
      {"dagger", 1, 6, 2, 7}, {"two-handed-sword", 4, 4, 1, 2}, {"dart", 1, 1, 8, 17}, {"shuriken", 1, 2, 8, 17},
      {"spear", 2, 3, 8, 17},
  };
  return T;
}
struct ArmorRow {
  std::string name;
  int def;
};
// BUILTIN_ARMORS (armor.rs:168-217)
const std::vector<ArmorRow>& builtin_armors() {
  static const std::vector<ArmorRow> T = {
      {"leather armor", 2}, {"ring mail", 3},   {"studded leather armor", 3}, {"scale mail", 4},
      {"chain mail", 5},    {"splint mail", 6}, {"banded mail", 6},           {"plate mail", 7},
  };
  return T;
}

struct Fail {
  std::string msg;
};
[[noreturn]] void fail(const std::string& m) { throw Fail{m}; }

const Value* field(const Value* o, const char* k) {
  if (!o) return nullptr;
  if (o->kind != Value::Object) fail(std::string("invalid type: expected a map for the parent of `") + k + "`");
  return o->get(k);
}
int64_t get_i64(const Value* o, const char* k, int64_t dflt, int64_t lo, int64_t hi) {
  const Value* v = field(o, k);
  if (!v) return dflt;
  if (!v->is_int()) fail(std::string("invalid type for `") + k + "`: expected an integer");
  int64_t x;
  try {
    x = v->as_i64();
  } catch (const std::exception&) {
    fail(std::string("invalid value for `") + k + "`: out of range");
  }
  if (x < lo || x > hi) fail(std::string("invalid value for `") + k + "`: out of range");
  return x;
}
uint32_t get_u32(const Value* o, const char* k, uint32_t dflt) { return (uint32_t)get_i64(o, k, dflt, 0, UINT32_MAX); }
int32_t get_i32(const Value* o, const char* k, int32_t dflt) { return (int32_t)get_i64(o, k, dflt, INT32_MIN, INT32_MAX); }

// symbol::Symbol::from_tile (symbol.rs:17-40)
int symbol_of(int t) {
  static const char table[] = " @#.-%+^!?])/*:=,";
  if (t == '|') return 4;
  for (int i = 0; table[i]; ++i)
    if (table[i] == t) return i;
  if (t >= 'A' && t <= 'Z') return t - 'A' + 17;
  return -1;
}

void u128_parts(const Value& v, const char* what, uint64_t* lo, uint64_t* hi) {
  if (!v.is_int() || v.negative) fail(std::string("invalid type for `") + what + "`: expected u128");
  *lo = (uint64_t)v.mag;
  *hi = (uint64_t)(v.mag >> 64);
}

void parse_into(const std::string& json, rg_params* p) {
  rgjson::ValuePtr rootp;
  try {
    rootp = rgjson::parse(json);
  } catch (const std::exception& e) {
    fail(e.what());
  }
  const Value* root = rootp.get();
  if (root->kind != Value::Object) fail("invalid type: expected struct GameConfig");
  memset(p, 0, sizeof(*p));
  p->width = get_i32(root, "width", 80);
  p->height = get_i32(root, "height", 24);
  if (const Value* s = field(root, "seed"); s && s->kind != Value::Null) {
    p->has_seed = 1;
    u128_parts(*s, "seed", &p->seed_lo, &p->seed_hi);
  }
  if (const Value* s = field(root, "seed_range"); s && s->kind != Value::Null) {
    if (s->kind != Value::Array || s->arr.size() != 2) fail("invalid type for `seed_range`: expected an array of length 2");
    uint64_t h0, h1;
    u128_parts(*s->arr[0], "seed_range", &p->seed_range_lo, &h0);
    u128_parts(*s->arr[1], "seed_range", &p->seed_range_hi, &h1);
    if (h0 || h1) fail("seed_range beyond 64 bits is not supported");
    p->has_seed_range = 1;
  }
  // ---- dungeon (DungeonStyle is internally tagged by "style", dungeon/mod.rs:16-28)
  const Value* dg = field(root, "dungeon");
  if (dg) {
    const Value* st = field(dg, "style");
    if (!st || st->kind != Value::String) fail("missing field `style`");
    if (st->str != "rogue") fail("unknown or unimplemented dungeon style `" + st->str + "`");
  }
  p->room_num_x = get_i32(dg, "room_num_x", 3);
  p->room_num_y = get_i32(dg, "room_num_y", 3);
  const Value* mrs = field(dg, "min_room_size");
  p->min_room_x = 4;
  p->min_room_y = 4;
  if (mrs) {
    if (!field(mrs, "x") || !field(mrs, "y")) fail("missing field in `min_room_size`");
    p->min_room_x = get_i32(mrs, "x", 4);
    p->min_room_y = get_i32(mrs, "y", 4);
  }
  p->max_empty_rooms = get_u32(dg, "max_empty_rooms", 3);
  p->amulet_level = get_u32(dg, "amulet_level", 25);
  p->maze_rate_inv = get_u32(dg, "maze_rate_inv", 15);
  p->dark_level = get_u32(dg, "dark_level", 10);
  p->hidden_passage_rate_inv = get_u32(dg, "hidden_passage_rate_inv", 40);
  p->locked_door_rate_inv = get_u32(dg, "locked_door_rate_inv", 5);
  p->max_extra_edges = get_u32(dg, "max_extra_edges", 5);
  p->door_unlock_rate_inv = get_u32(dg, "door_unlock_rate_inv", 5);
  p->passage_unlock_rate_inv = get_u32(dg, "passage_unlock_rate_inv", 3);
  // ---- items
  const Value* item = field(root, "item");
  const Value* gold = field(item, "gold");
  p->gold_rate_inv = get_u32(gold, "rate_inv", 2);
  p->gold_base = get_u32(gold, "base", 50);
  p->gold_per_level = get_u32(gold, "per_level", 10);
  p->gold_minimum = get_u32(gold, "minimum", 2);
  std::vector<WeaponRow> weapons;
  if (const Value* ws = field(field(item, "weapon"), "weapons")) {
    if (ws->kind != Value::Array) fail("invalid type for `weapons`: expected a sequence");
    for (const auto& w : ws->arr) {
      if (w->is_int()) {
        uint64_t i = w->as_u64();
        if (i >= builtin_weapons().size()) fail("weapon preset index out of range");
        weapons.push_back(builtin_weapons()[i]);
      } else if (w->kind == Value::Object) {
        const Value* nm = field(w.get(), "name");
        const Value* aw = field(w.get(), "at_weild");
        const Value* in = field(w.get(), "init_num");
        if (!nm || nm->kind != Value::String || !aw || !in) fail("data did not match any variant of untagged enum Preset");
        weapons.push_back(WeaponRow{nm->str, get_i32(aw, "times", 0), get_i32(aw, "max", 0), get_u32(in, "start", 0),
                                    get_u32(in, "end", 0)});
      } else {
        fail("data did not match any variant of untagged enum Preset");
      }
    }
  } else {
    weapons = builtin_weapons();
  }
  std::vector<ArmorRow> armors;
  if (const Value* as = field(field(item, "armor"), "armors")) {
    if (as->kind != Value::Array) fail("invalid type for `armors`: expected a sequence");
    for (const auto& a : as->arr) {
      if (a->is_int()) {
        uint64_t i = a->as_u64();
        if (i >= builtin_armors().size()) fail("armor preset index out of range");
        armors.push_back(builtin_armors()[i]);
      } else if (a->kind == Value::Object) {
        const Value* nm = field(a.get(), "name");
        if (!nm || nm->kind != Value::String || !field(a.get(), "def")) fail("data did not match any variant of untagged enum Preset");
        armors.push_back(ArmorRow{nm->str, get_i32(a.get(), "def", 0)});
      } else {
        fail("data did not match any variant of untagged enum Preset");
      }
    }
  } else {
    armors = builtin_armors();
  }
  // ---- player
  const Value* pl = field(root, "player");
  static const uint32_t DEFAULT_EXPS[21] = {10,    20,    40,     80,     160,    320,    640,     1300,    2600,    5200,      13000,
                                            26000, 50000, 100000, 200000, 400000, 800000, 2000000, 4000000, 8000000, 0xFFFFFFFFu};
  if (const Value* ex = field(pl, "exps")) {
    if (ex->kind != Value::Array) fail("invalid type for `exps`: expected a sequence");
    if (ex->arr.size() > RG_MAX_EXPS) fail("more than 32 experience thresholds are not supported");
    p->n_exps = (uint32_t)ex->arr.size();
    for (size_t i = 0; i < ex->arr.size(); ++i) {
      if (!ex->arr[i]->is_int()) fail("invalid type in `exps`: expected u32");
      uint64_t v = ex->arr[i]->as_u64();
      if (v > UINT32_MAX) fail("invalid value in `exps`: expected u32");
      p->exps[i] = (uint32_t)v;
    }
  } else {
    p->n_exps = 21;
    for (int i = 0; i < 21; ++i) p->exps[i] = DEFAULT_EXPS[i];
  }
  p->hunger_time = get_u32(pl, "hunger_time", 1300);
  p->init_hp = (int32_t)get_i64(pl, "init_hp", 12, INT32_MIN, INT32_MAX);
  const int64_t max_items = get_i64(pl, "max_items", 27, 0, INT64_MAX);
  // init_items (item/mod.rs:160-180). Default = Player default_init_items player.rs:66-75.
  struct Init {
    int kind;  // 0 noinit-gold, 1 noinit-other, 2 armor, 3 weapon
    std::string name;
    uint32_t how_many;
    int a, b;  // def_plus | hit_plus, dam_plus
  };
  std::vector<Init> inits;
  if (const Value* ii = field(pl, "init_items")) {
    if (ii->kind != Value::Array) fail("invalid type for `init_items`: expected a sequence");
    for (const auto& it : ii->arr) {
      if (it->kind != Value::Object || it->obj.size() != 1) fail("invalid InitItem: expected a single-key map");
      const std::string& tag = it->obj[0].first;
      const Value* body = it->obj[0].second.get();
      if (tag == "Noinit") {
        const Value* kind = field(body, "kind");
        if (!kind) fail("missing field `kind`");
        bool is_gold = kind->kind == Value::String && kind->str == "Gold";
        inits.push_back(Init{is_gold ? 0 : 1, "", get_u32(body, "how_many", 0), 0, 0});
      } else if (tag == "Armor") {
        const Value* nm = field(body, "name");
        if (!nm || nm->kind != Value::String) fail("missing field `name`");
        inits.push_back(Init{2, nm->str, 0, get_i32(body, "def_plus", 0), 0});
      } else if (tag == "Weapon") {
        const Value* nm = field(body, "name");
        if (!nm || nm->kind != Value::String) fail("missing field `name`");
        inits.push_back(Init{3, nm->str, get_u32(body, "num_plus", 0), get_i32(body, "hit_plus", 0), get_i32(body, "dam_plus", 0)});
      } else {
        fail("unknown variant `" + tag + "`, expected one of `Noinit`, `Armor`, `Weapon`");
      }
    }
  } else {
    inits = {{0, "", 0, 0, 0},        {1, "", 1, 0, 0},       {2, "ring mail", 0, 1, 0},
             {3, "mace", 0, 1, 1},    {3, "bow", 0, 1, 0},    {3, "arrow", 25, 0, 0}};
  }
  bool has_gold_stack = false, have_weapon = false, have_armor = false;
  p->weapon_times = 1;  // fight.rs:31: bare hands roll 1d4
  p->weapon_max = 4;
  for (const Init& it : inits) {
    if (it.kind == 0 && !has_gold_stack) {
      has_gold_stack = true;
      p->init_gold = it.how_many;
    } else if (it.kind == 3) {
      const WeaponRow* row = nullptr;
      for (const auto& w : weapons)
        if (w.name == it.name) { row = &w; break; }
      if (!row) fail("Specified item " + it.name + " is not registerd to WeaponHandler");
      if (p->n_init_draws >= RG_MAX_INIT_DRAWS) fail("more than 8 initial weapons are not supported");
      p->init_draw_lo[p->n_init_draws] = row->num_lo;
      p->init_draw_hi[p->n_init_draws] = row->num_hi;
      p->n_init_draws += 1;
      if (!have_weapon) {
        have_weapon = true;
        p->weapon_times = row->times;
        p->weapon_max = row->max;
        p->weapon_hit_plus = it.a;
        p->weapon_dam_plus = it.b;
      }
    } else if (it.kind == 2) {
      const ArmorRow* row = nullptr;
      for (const auto& a : armors)
        if (a.name == it.name) { row = &a; break; }
      if (!row) fail("Specified item " + it.name + " is not registerd to WeaponHandler");
      if (!have_armor) {
        have_armor = true;
        p->armor_def = row->def + it.a;
      }
    }
  }
  p->pack_accepts_gold = (has_gold_stack || (int64_t)inits.size() < max_items) ? 1 : 0;
  // ---- monsters
  const Value* en = field(root, "enemies");
  std::vector<KindRow> kinds;
  if (const Value* es = field(en, "enemies")) {
    if (es->kind != Value::Array) fail("invalid type for `enemies`: expected a sequence");
    for (const auto& e : es->arr) {
      if (e->is_int()) {
        uint64_t i = e->as_u64();
        if (i >= 26) fail("enemy preset index out of range");
        kinds.push_back(builtin_enemy(i));
      } else if (e->kind == Value::Object) {
        KindRow k;
        const Value* at = field(e.get(), "attack");
        if (!at || at->kind != Value::Array || !field(e.get(), "tile") || !field(e.get(), "level"))
          fail("data did not match any variant of untagged enum Preset");
        for (const auto& d : at->arr) k.dice.push_back({get_i32(d.get(), "times", 0), get_i32(d.get(), "max", 0)});
        k.attr = get_u32(e.get(), "attr", 0);
        k.defense = get_i32(e.get(), "defense", 0);
        k.exp = get_u32(e.get(), "exp", 0);
        k.level = get_i32(e.get(), "level", 1);
        k.rarelity = get_i32(e.get(), "rarelity", 0);
        k.tile = get_i32(e.get(), "tile", 'A');
        kinds.push_back(k);
      } else {
        fail("data did not match any variant of untagged enum Preset");
      }
    }
  } else {
    for (size_t i = 0; i < 26; ++i) kinds.push_back(builtin_enemy(i));
  }
  if (kinds.size() > RG_MAX_ENEMY_KINDS) fail("more than 32 monster kinds are not supported");
  int tile_max = -1;
  for (const auto& k : kinds) tile_max = std::max(tile_max, k.tile);
  std::stable_sort(kinds.begin(), kinds.end(), [](const KindRow& a, const KindRow& b) { return a.rarelity < b.rarelity; });
  p->n_enemies = (uint32_t)kinds.size();
  for (size_t i = 0; i < kinds.size(); ++i) {
    rg_enemy_kind& o = p->enemies[i];
    const KindRow& k = kinds[i];
    if (k.dice.size() > RG_MAX_DICE) fail("more than 4 attack dice are not supported");
    o.tile = k.tile;
    o.level = k.level;
    o.defense = k.defense;
    o.exp = k.exp;
    o.attr = k.attr;
    o.n_dice = (uint32_t)k.dice.size();
    for (size_t j = 0; j < k.dice.size(); ++j) {
      o.dice_times[j] = k.dice[j].first;
      o.dice_max[j] = k.dice[j].second;
    }
  }
  p->appear_rate_gold = get_u32(en, "appear_rate_gold", 80);
  p->appear_rate_nogold = get_u32(en, "appear_rate_nogold", 25);
  if (const Value* hd = field(root, "hide_dungeon")) {
    if (hd->kind != Value::Bool) fail("invalid type for `hide_dungeon`: expected a boolean");
    p->hide_dungeon = hd->b ? 1 : 0;
  } else {
    p->hide_dungeon = 1;
  }
  // GameConfig::symbol_max (+1: state_impls.rs:21-25)
  int sym = kinds.empty() ? symbol_of('A') - 1 : symbol_of(tile_max);
  if (sym < 0) fail("Failed to get symbol max");
  p->symbols = (uint32_t)sym + 1;
}

void put_err(char* err, size_t n, const std::string& m) {
  if (err && n) snprintf(err, n, "%s", m.c_str());
}

}  // namespace

extern "C" int rg_parse_config(const char* json, rg_params* out, char* err, size_t err_len) {
  if (!json || !out) {
    put_err(err, err_len, "null argument");
    return RG_ERR_ARG;
  }
  try {
    parse_into(json, out);
  } catch (const Fail& f) {
    put_err(err, err_len, "Failed to parse config: " + f.msg);
    return RG_ERR_PARSE;
  } catch (const std::exception& e) {
    put_err(err, err_len, std::string("Failed to parse config: ") + e.what());
    return RG_ERR_PARSE;
  }
  return RG_OK;
}

// to_global's size checks (core/src/lib.rs:166-184) plus the conditions under which the
// reference's generator panics or this implementation's fixed-size tables do not fit.
extern "C" int rg_validate_params(const rg_params* p, char* err, size_t err_len) {
  auto bad = [&](const char* m) {
    put_err(err, err_len, std::string("Error in rogue-gym: invalid setting: ") + m);
    return (int)RG_ERR_SETTING;
  };
  if (p->width < 32) return bad("screen width is too narrow");
  if (p->width > 160) return bad("screen width is too wide");
  if (p->height < 16) return bad("screen height is too narrow");
  if (p->height > 48) return bad("screen height is too wide");
  if (p->room_num_x < 1 || p->room_num_y < 1) return bad("room_num must be positive");
  if (p->room_num_x * p->room_num_y > RG_MAX_ROOMS) return bad("more than 16 rooms per floor are not supported");
  // rooms.rs:256-259: range(min_room_size..room_size) panics unless min < size on both axes;
  // the tightest sector is a top/bottom-row one that lost a row (rooms.rs:196-206).
  const int rsx = p->width / p->room_num_x;
  int rsy = p->height / p->room_num_y - 1;
  if (p->room_num_y == 1) rsy -= 1;
  if (p->min_room_x < 3 || p->min_room_y < 3) return bad("min_room_size must be at least 3x3");
  if (p->min_room_x >= rsx || p->min_room_y >= rsy) return bad("room grid does not fit the screen (min_room_size >= sector size)");
  if (p->dark_level == 0 || p->maze_rate_inv == 0 || p->hidden_passage_rate_inv == 0 || p->locked_door_rate_inv == 0 ||
      p->max_extra_edges == 0 || p->door_unlock_rate_inv == 0 || p->passage_unlock_rate_inv == 0 || p->gold_rate_inv == 0)
    return bad("a rate/level parameter is 0 (the reference panics in RngHandle::range)");
  if (p->gold_base + p->gold_per_level == 0) return bad("gold base + per_level is 0");
  for (uint32_t i = 0; i < p->n_init_draws; ++i)
    if (p->init_draw_lo[i] >= p->init_draw_hi[i]) return bad("empty init_num range");
  for (uint32_t i = 0; i < p->n_enemies; ++i) {
    if (p->enemies[i].level < 1) return bad("monster level must be >= 1");
    for (uint32_t j = 0; j < p->enemies[i].n_dice; ++j)
      if (p->enemies[i].dice_times[j] > 0 && p->enemies[i].dice_max[j] < 1) return bad("monster attack dice max must be >= 1");
  }
  if (p->weapon_times > 0 && p->weapon_max < 1) return bad("weapon dice max must be >= 1");
  if (p->init_hp < 1) return bad("init_hp must be positive");
  return RG_OK;
}
