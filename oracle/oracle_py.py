"""ctypes binding of the CPU oracle (oracle/oracle.h) + an independent config parser.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package never imports this module.

The JSON -> flat-parameter conversion below is written independently of the product's C++
config parser (rogue-gym_b200/csrc/config.cpp) so that parity tests also cover parsing.
Reference schema: core/src/lib.rs:42-86, dungeon/rogue/mod.rs:22-61, character/player.rs:16-32,
character/enemies.rs:17-27, item/{gold.rs:7-16,weapon.rs,armor.rs}.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")

MAX_ENEMY_KINDS, MAX_DICE, MAX_EXPS, MAX_INIT_DRAWS, MAX_ROOMS, DIST_CACHE = 32, 4, 32, 8, 16, 9


class EnemyKind(C.Structure):
    _fields_ = [
        ("tile", C.c_int32), ("level", C.c_int32), ("defense", C.c_int32), ("exp", C.c_uint32),
        ("attr", C.c_uint32), ("n_dice", C.c_uint32),
        ("dice_times", C.c_int32 * MAX_DICE), ("dice_max", C.c_int32 * MAX_DICE),
    ]


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("room_num_x", C.c_int32), ("room_num_y", C.c_int32), ("min_room_x", C.c_int32), ("min_room_y", C.c_int32),
        ("max_empty_rooms", C.c_uint32), ("amulet_level", C.c_uint32), ("maze_rate_inv", C.c_uint32),
        ("dark_level", C.c_uint32), ("hidden_passage_rate_inv", C.c_uint32), ("locked_door_rate_inv", C.c_uint32),
        ("max_extra_edges", C.c_uint32), ("door_unlock_rate_inv", C.c_uint32), ("passage_unlock_rate_inv", C.c_uint32),
        ("gold_rate_inv", C.c_uint32), ("gold_base", C.c_uint32), ("gold_per_level", C.c_uint32),
        ("gold_minimum", C.c_uint32),
        ("hunger_time", C.c_uint32), ("init_hp", C.c_int32), ("n_exps", C.c_uint32), ("exps", C.c_uint32 * MAX_EXPS),
        ("pack_accepts_gold", C.c_int32), ("init_gold", C.c_uint32),
        ("weapon_times", C.c_int32), ("weapon_max", C.c_int32), ("weapon_hit_plus", C.c_int32),
        ("weapon_dam_plus", C.c_int32), ("armor_def", C.c_int32),
        ("n_init_draws", C.c_uint32), ("init_draw_lo", C.c_uint32 * MAX_INIT_DRAWS),
        ("init_draw_hi", C.c_uint32 * MAX_INIT_DRAWS),
        ("n_enemies", C.c_uint32), ("enemies", EnemyKind * MAX_ENEMY_KINDS),
        ("appear_rate_gold", C.c_uint32), ("appear_rate_nogold", C.c_uint32),
        ("hide_dungeon", C.c_int32), ("symbols", C.c_uint32),
    ]


class Scalars(C.Structure):
    _fields_ = [
        ("level", C.c_int32), ("px", C.c_int32), ("py", C.c_int32), ("hp", C.c_int32), ("hp_max", C.c_int32),
        ("exp", C.c_uint32), ("plevel", C.c_int32), ("food_left", C.c_uint32), ("quiet", C.c_uint32),
        ("gold", C.c_uint32), ("ui_dead", C.c_int32), ("steps", C.c_int32), ("is_terminal", C.c_int32),
        ("message", C.c_uint32), ("error", C.c_int32), ("n_monsters", C.c_int32), ("n_items", C.c_int32),
        ("n_cache", C.c_int32), ("status", C.c_uint32 * 10), ("rng", C.c_uint32 * 12),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


# EnemyAttr bits, core/src/character/enemies.rs:125-137
MEAN, FLYING, REGENERATE, GREEDY, INVISIBLE, RUSTS_ARMOR, STEAL_GOLD, REDUCE_STR, FREEZES, RANDOM, CONFUSED = (
    1 << i for i in range(11)
)

# BUILTIN_ENEMIES, core/src/character/enemies.rs:474-761, in declaration order (index = preset id):
# (tile, attack dice [(times, max)], attr, defense, exp, level, rarelity)
BUILTIN_ENEMIES = [
    ("A", [(0, 0)], MEAN | RUSTS_ARMOR, 2 | 8, 20, 5, 12),
    ("B", [(1, 2)], FLYING | RANDOM, 3, 1, 1, 2),
    ("C", [(1, 2), (1, 5), (1, 5)], 0, 4, 17, 4, 10),
    ("D", [(1, 8), (1, 8), (3, 10)], MEAN, 3, 5000, 10, 25),
    ("E", [(1, 2)], MEAN, 7, 2, 1, 1),
    ("F", [], MEAN, 3, 80, 8, 15),
    ("G", [(4, 3), (3, 5)], FLYING | MEAN | REGENERATE, 2, 2000, 13, 23),
    ("H", [(1, 8)], MEAN, 5, 3, 1, 4),
    ("I", [(0, 0)], FREEZES, 9, 5, 1, 5),
    ("J", [(2, 12), (2, 4)], 0, 6, 3000, 15, 24),
    ("K", [(1, 4)], MEAN, 7, 1, 1, 0),
    ("L", [(1, 1)], STEAL_GOLD, 8, 10, 3, 9),
    ("M", [(3, 4), (3, 4), (2, 5)], MEAN, 2, 200, 8, 21),
    ("N", [(0, 0)], 0, 9, 37, 3, 13),
    ("O", [(1, 8)], GREEDY, 6, 5, 1, 7),
    ("P", [(4, 4)], INVISIBLE, 3, 120, 8, 18),
    ("Q", [(1, 5), (1, 5)], MEAN, 3, 15, 3, 11),
    ("R", [(1, 6)], REDUCE_STR | MEAN, 3, 9, 2, 6),
    ("S", [(1, 3)], MEAN, 5, 2, 1, 3),
    ("T", [(1, 8), (1, 8), (2, 6)], MEAN | REGENERATE, 4, 120, 6, 16),
    ("U", [(1, 9), (1, 9), (2, 9)], MEAN, -2, 190, 7, 20),
    ("V", [(1, 19)], MEAN | REGENERATE, 1, 350, 8, 22),
    ("W", [(1, 6)], 0, 4, 55, 5, 17),
    ("X", [(4, 4)], 0, 7, 100, 7, 19),
    ("Y", [(1, 6), (1, 6)], 0, 6, 50, 4, 14),
    ("Z", [(1, 8)], MEAN, 8, 6, 2, 8),
]

# BUILTIN_WEAPONS core/src/item/weapon.rs:196-296: name -> (at_weild (times,max), init_num lo..hi)
BUILTIN_WEAPONS = [
    ("mace", (2, 4), (1, 2)),
    ("long-sword", (3, 4), (1, 2)),
    ("bow", (1, 1), (1, 2)),
    ("arrow", (1, 1), (8, 17)),
    ("dagger", (1, 6), (2, 7)),
    ("two-handed-sword", (4, 4), (1, 2)),
    ("dart", (1, 1), (8, 17)),
    ("shuriken", (1, 2), (8, 17)),
    ("spear", (2, 3), (8, 17)),
]
# BUILTIN_ARMORS core/src/item/armor.rs:168-217: name -> def
BUILTIN_ARMORS = [
    ("leather armor", 2), ("ring mail", 3), ("studded leather armor", 3), ("scale mail", 4),
    ("chain mail", 5), ("splint mail", 6), ("banded mail", 6), ("plate mail", 7),
]

DEFAULT_EXPS = [10, 20, 40, 80, 160, 320, 640, 1300, 2600, 5200, 13000, 26000, 50000, 100000, 200000,
                400000, 800000, 2000000, 4000000, 8000000, 0xFFFFFFFF]
DEFAULT_INIT_ITEMS = [
    {"Noinit": {"kind": "Gold", "how_many": 0, "attr": 4}},
    {"Noinit": {"kind": {"Food": "Ration"}, "how_many": 1, "attr": 4}},
    {"Armor": {"name": "ring mail", "def_plus": 1}},
    {"Weapon": {"name": "mace", "num_plus": 0, "hit_plus": 1, "dam_plus": 1}},
    {"Weapon": {"name": "bow", "num_plus": 0, "hit_plus": 1, "dam_plus": 0}},
    {"Weapon": {"name": "arrow", "num_plus": 25, "hit_plus": 0, "dam_plus": 0}},
]


def symbol_of_tile(ch):
    table = " @#.-%+^!?])/*:=,"
    if ch == "|":
        return 4
    if ch in table:
        return table.index(ch)
    if "A" <= ch <= "Z":
        return ord(ch) - ord("A") + 17
    return None


def params_from_config(cfg):
    """GameConfig JSON (dict or str) -> (Params, seed or None)."""
    if isinstance(cfg, str):
        cfg = json.loads(cfg)
    p = Params()
    p.width = cfg.get("width", 80)
    p.height = cfg.get("height", 24)
    dg = cfg.get("dungeon", {"style": "rogue"})
    if dg.get("style", "rogue") != "rogue":
        raise ValueError("only the rogue dungeon style exists")
    p.room_num_x = dg.get("room_num_x", 3)
    p.room_num_y = dg.get("room_num_y", 3)
    mrs = dg.get("min_room_size", {"x": 4, "y": 4})
    p.min_room_x, p.min_room_y = mrs["x"], mrs["y"]
    p.max_empty_rooms = dg.get("max_empty_rooms", 3)
    p.amulet_level = dg.get("amulet_level", 25)
    p.maze_rate_inv = dg.get("maze_rate_inv", 15)
    p.dark_level = dg.get("dark_level", 10)
    p.hidden_passage_rate_inv = dg.get("hidden_passage_rate_inv", 40)
    p.locked_door_rate_inv = dg.get("locked_door_rate_inv", 5)
    p.max_extra_edges = dg.get("max_extra_edges", 5)
    p.door_unlock_rate_inv = dg.get("door_unlock_rate_inv", 5)
    p.passage_unlock_rate_inv = dg.get("passage_unlock_rate_inv", 3)
    item = cfg.get("item", {})
    gold = item.get("gold", {})
    p.gold_rate_inv = gold.get("rate_inv", 2)
    p.gold_base = gold.get("base", 50)
    p.gold_per_level = gold.get("per_level", 10)
    p.gold_minimum = gold.get("minimum", 2)
    weapons = []
    for w in item.get("weapon", {}).get("weapons", list(range(len(BUILTIN_WEAPONS)))):
        if isinstance(w, int):
            weapons.append(BUILTIN_WEAPONS[w])
        else:
            weapons.append((w["name"], (w["at_weild"]["times"], w["at_weild"]["max"]),
                            (w["init_num"]["start"], w["init_num"]["end"])))
    armors = []
    for a in item.get("armor", {}).get("armors", list(range(len(BUILTIN_ARMORS)))):
        armors.append(BUILTIN_ARMORS[a] if isinstance(a, int) else (a["name"], a["def"]))
    pl = cfg.get("player", {})
    exps = pl.get("exps", DEFAULT_EXPS)
    p.n_exps = len(exps)
    for i, e in enumerate(exps):
        p.exps[i] = e
    p.hunger_time = pl.get("hunger_time", 1300)
    p.init_hp = pl.get("init_hp", 12)
    max_items = pl.get("max_items", 27)
    init_items = pl.get("init_items", DEFAULT_INIT_ITEMS)
    has_gold_stack, init_gold = False, 0
    weapon, armor, draws = None, None, []
    for it in init_items:
        if "Noinit" in it:
            if it["Noinit"]["kind"] == "Gold" and not has_gold_stack:
                has_gold_stack, init_gold = True, it["Noinit"]["how_many"]
        elif "Weapon" in it:
            w = it["Weapon"]
            st = next(s for s in weapons if s[0] == w["name"])
            draws.append(st[2])
            if weapon is None:
                weapon = (st[1][0], st[1][1], w["hit_plus"], w["dam_plus"])
        elif "Armor" in it:
            a = it["Armor"]
            st = next(s for s in armors if s[0] == a["name"])
            if armor is None:
                armor = st[1] + a["def_plus"]
    p.pack_accepts_gold = 1 if (has_gold_stack or len(init_items) < max_items) else 0
    p.init_gold = init_gold
    if weapon is None:
        weapon = (1, 4, 0, 0)  # fight.rs:31 default dice, no plus
    p.weapon_times, p.weapon_max, p.weapon_hit_plus, p.weapon_dam_plus = weapon
    p.armor_def = armor if armor is not None else 0
    p.n_init_draws = len(draws)
    for i, (lo, hi) in enumerate(draws):
        p.init_draw_lo[i], p.init_draw_hi[i] = lo, hi
    en = cfg.get("enemies", {})
    kinds = []
    for preset in en.get("enemies", list(range(26))):
        if isinstance(preset, int):
            kinds.append(BUILTIN_ENEMIES[preset])
        else:
            kinds.append((chr(preset["tile"]) if isinstance(preset["tile"], int) else preset["tile"],
                          [(d["times"], d["max"]) for d in preset["attack"]], preset["attr"],
                          preset["defense"], preset["exp"], preset["level"], preset["rarelity"]))
    kinds = sorted(kinds, key=lambda k: k[6])  # stable, enemies.rs:251
    p.n_enemies = len(kinds)
    for i, (tile, dice, attr, defense, exp, level, _r) in enumerate(kinds):
        k = p.enemies[i]
        k.tile, k.level, k.defense, k.exp, k.attr, k.n_dice = ord(tile), level, defense, exp, attr, len(dice)
        for j, (t, m) in enumerate(dice):
            k.dice_times[j], k.dice_max[j] = t, m
    p.appear_rate_gold = en.get("appear_rate_gold", 80)
    p.appear_rate_nogold = en.get("appear_rate_nogold", 25)
    p.hide_dungeon = 1 if cfg.get("hide_dungeon", True) else 0
    # GameConfig::symbol_max core/src/lib.rs:150-155, +1 (state_impls.rs:21-25)
    if kinds:
        p.symbols = symbol_of_tile(max(k[0] for k in kinds)) + 1
    else:
        p.symbols = symbol_of_tile("A") - 1 + 1
    return p, cfg.get("seed", None)


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
        os.path.getmtime(os.path.join(HERE, f)) for f in ("oracle.cpp", "oracle.h")
    ):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, u8p = C.c_void_p, C.POINTER(C.c_uint8)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(Params), C.c_int64]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_seed.argtypes = [vp, C.c_uint64, C.c_uint64]
        L.orc_reset.argtypes = [vp]
        L.orc_react.argtypes = [vp, C.c_uint8]
        L.orc_step_auto.argtypes = [vp, C.c_uint8]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [vp]
        L.orc_get_obs.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_get_scalars.argtypes = [vp, C.POINTER(Scalars)]
        L.orc_get_grid.argtypes = [vp, vp, vp]
        L.orc_get_entities.argtypes = [vp, vp, vp]
        L.orc_get_dist_cache.argtypes = [vp, vp, vp]
        L.orc_get_rooms.argtypes = [vp, vp]
        L.orc_get_draw_counts.argtypes = [vp, vp]
        L.orc_encode.argtypes = [vp, C.c_int, C.c_uint32, C.c_int, vp]
        L.orc_test_move_enemy.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_state_hash.restype = C.c_uint64
        L.orc_state_hash.argtypes = [vp]
        L.orc_batch_rollout.restype = C.c_double
        L.orc_batch_rollout.argtypes = [C.POINTER(vp), C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                        C.POINTER(C.c_uint64)]
        L.orc_batch_step.argtypes = [C.POINTER(vp), C.c_int64, vp, C.c_int, C.c_int, vp]
        L.orc_batch_reset.argtypes = [C.POINTER(vp), C.c_int64, C.c_int, vp]
        L.orc_batch_get_obs.argtypes = [C.POINTER(vp), C.c_int64, vp, vp, vp, vp, vp]
        L.orc_batch_hash.argtypes = [C.POINTER(vp), C.c_int64, vp]
        L.orc_test_ordset.restype = C.c_int64
        L.orc_test_ordset.argtypes = [C.c_uint64, vp, vp, vp, C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]
        L.orc_test_ordset_from_range_contains.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_test_edges.argtypes = [C.c_int] * 6 + [vp, C.c_int]
        L.orc_test_dice.restype = None
        L.orc_test_dice.argtypes = [C.c_uint64, C.c_int, C.c_int64, C.c_int64, vp]
        L.orc_test_select_cell.restype = C.c_int64
        L.orc_test_select_cell.argtypes = [C.POINTER(Params), C.c_uint32, C.c_uint64, C.c_int64]
        _lib = L
    return _lib


def ordset_nth(cap, members, k, ops=None):
    """FenwickSet restatement: apply insert (1) / remove (2) ops for `members`, return (nth(k) or None, len, results)."""
    m = np.asarray(members, np.uint64)
    o = np.ones(len(m), np.uint8) if ops is None else np.asarray(ops, np.uint8)
    res = np.zeros(len(m), np.uint8)
    n_out = C.c_uint64()
    v = lib().orc_test_ordset(cap, m.ctypes.data, o.ctypes.data, res.ctypes.data, len(m), k, C.byref(n_out))
    return (None if v < 0 else int(v)), int(n_out.value), res.astype(bool)


def edges(x0, x1, y0, y1, direction, inclusive):
    """passages::edges of a half-open rect; direction 0 Up, 1 Down, 2 Left, 3 Right. List of (x, y)."""
    out = np.zeros((512, 2), np.int32)
    n = lib().orc_test_edges(x0, x1, y0, y1, direction, int(inclusive), out.ctypes.data, 512)
    return [tuple(int(v) for v in p) for p in out[:n]]


def dice(seed, times, mx, count):
    """`count` rolls of `times` d `mx` (Dice::random, character/mod.rs:229-235) on a stream seeded with `seed`."""
    out = np.zeros(count, np.int64)
    lib().orc_test_dice(seed, times, mx, count, out.ctypes.data)
    return out


def select_cells(config, level, seed, tries=1000):
    """floor.rs:490-505: cells Floor::select_cell hands out on a fresh floor of `level` before it runs dry (-1: set_obj refused)."""
    params, _ = params_from_config(config)
    return int(lib().orc_test_select_cell(C.byref(params), level, seed, tries))


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("oracle error %d: %s" % (code, msg))
        self.code = code


class OracleEnv:
    """One GameStateImpl (python/src/state_impls.rs) on the CPU oracle."""

    def __init__(self, config, max_steps=1000, seed=None, reset=True):
        self.L = lib()
        self.params, cfg_seed = params_from_config(config)
        self.w, self.h = self.params.width, self.params.height
        self.C = self.w * self.h
        self.ptr = self.L.orc_create(C.byref(self.params), max_steps)
        s = seed if seed is not None else cfg_seed
        if s is not None:
            self.set_seed(s)
        if reset:
            self.reset()

    def __del__(self):
        if getattr(self, "ptr", None):
            self.L.orc_destroy(self.ptr)
            self.ptr = None

    def set_seed(self, seed):
        self.L.orc_set_seed(self.ptr, seed & 0xFFFFFFFFFFFFFFFF, (seed >> 64) & 0xFFFFFFFFFFFFFFFF)

    def _check(self, rc):
        if rc != 0:
            raise OracleError(rc, self.L.orc_last_error(self.ptr).decode())

    def reset(self):
        self._check(self.L.orc_reset(self.ptr))

    def react(self, key):
        self._check(self.L.orc_react(self.ptr, key if isinstance(key, int) else ord(key)))

    def react_str(self, keys):
        for k in keys:
            self.react(k)

    def step_auto(self, key):
        self._check(self.L.orc_step_auto(self.ptr, key if isinstance(key, int) else ord(key)))

    def obs(self):
        screen = np.zeros(self.C, np.uint8)
        hist = np.zeros(self.C, np.uint8)
        status = np.zeros(10, np.uint32)
        msg = C.c_uint32()
        term = C.c_int32()
        self.L.orc_get_obs(self.ptr, screen.ctypes.data, hist.ctypes.data, status.ctypes.data, C.addressof(msg),
                           C.addressof(term))
        return dict(screen=screen.reshape(self.h, self.w), history=hist.reshape(self.h, self.w), status=status,
                    message=msg.value, is_terminal=bool(term.value))

    def dungeon(self):
        return [bytes(r).decode() for r in self.obs()["screen"]]

    def scalars(self):
        s = Scalars()
        self.L.orc_get_scalars(self.ptr, C.byref(s))
        return s

    def grid(self):
        surface = np.zeros(self.C, np.uint8)
        attr = np.zeros(self.C, np.uint8)
        self.L.orc_get_grid(self.ptr, surface.ctypes.data, attr.ctypes.data)
        return surface.reshape(self.h, self.w), attr.reshape(self.h, self.w)

    def entities(self):
        m = np.zeros((MAX_ROOMS, 8), np.int32)
        it = np.zeros((MAX_ROOMS, 3), np.int32)
        self.L.orc_get_entities(self.ptr, m.ctypes.data, it.ctypes.data)
        s = self.scalars()
        return m[: s.n_monsters], it[: s.n_items]

    def dist_cache(self, with_maps=True):
        xy = np.zeros((DIST_CACHE, 2), np.int32)
        maps = np.zeros((DIST_CACHE, self.C), np.uint16) if with_maps else None
        self.L.orc_get_dist_cache(self.ptr, xy.ctypes.data, maps.ctypes.data if with_maps else None)
        n = self.scalars().n_cache
        return xy[:n], (maps[:n] if with_maps else None)

    def rooms(self):
        r = np.zeros((MAX_ROOMS, 8), np.int32)
        self.L.orc_get_rooms(self.ptr, r.ctypes.data)
        return r[: self.params.room_num_x * self.params.room_num_y]

    def draw_counts(self):
        c = np.zeros(3, np.uint64)
        self.L.orc_get_draw_counts(self.ptr, c.ctypes.data)
        return c

    def encode(self, mode, flag=0, with_hist=False):
        base = 1 if mode == 0 else self.params.symbols
        ch = base + bin(flag & 0x1FF).count("1") + (1 if with_hist else 0)
        out = np.zeros((ch, self.h, self.w), np.float32)
        rc = self.L.orc_encode(self.ptr, mode, flag, int(with_hist), out.ctypes.data)
        if rc < 0:
            raise OracleError(-1, "InvalidTileError")
        return out

    def test_move_enemy(self, frm, to):
        nx, ny = C.c_int(), C.c_int()
        k = self.L.orc_test_move_enemy(self.ptr, frm[0], frm[1], to[0], to[1], C.byref(nx), C.byref(ny))
        return k, (nx.value, ny.value)

    def state_hash(self):
        return self.L.orc_state_hash(self.ptr)


def batch_rollout(envs, first_env_id, t0, steps, threads):
    """Times `steps` auto-reset steps of the SURVEY §8d action stream over `envs` on the host."""
    L = lib()
    arr = (C.c_void_p * len(envs))(*[e.ptr for e in envs])
    dig = C.c_uint64()
    secs = L.orc_batch_rollout(arr, len(envs), first_env_id, t0, steps, threads, 0, C.byref(dig))
    return secs, dig.value


class OracleBatch:
    """N independent oracle envs stepped in lockstep (the CPU side of the parity tests and of
    bench.py's cpu_baseline). Seeds: env i gets seeds[i]."""

    def __init__(self, config, n, max_steps=1000, seeds=None, threads=None):
        self.L = lib()
        self.envs = [OracleEnv(config, max_steps, seed=(seeds[i] if seeds is not None else None), reset=False)
                     for i in range(n)]
        self.n = n
        self.w, self.h, self.C = self.envs[0].w, self.envs[0].h, self.envs[0].C
        self.ptrs = (C.c_void_p * n)(*[e.ptr for e in self.envs])
        self.threads = threads or min(os.cpu_count() or 1, 16)
        self.rc = np.zeros(n, np.int32)

    def seed(self, seeds):
        for e, s in zip(self.envs, seeds):
            e.set_seed(int(s))

    def reset(self):
        self.L.orc_batch_reset(self.ptrs, self.n, self.threads, self.rc.ctypes.data)
        return self.rc

    def step(self, keys, auto_reset=True):
        keys = np.ascontiguousarray(keys, np.uint8)
        self.L.orc_batch_step(self.ptrs, self.n, keys.ctypes.data, int(auto_reset), self.threads, self.rc.ctypes.data)
        return self.rc

    def obs(self):
        n, Cc = self.n, self.C
        screen = np.zeros((n, Cc), np.uint8)
        hist = np.zeros((n, Cc), np.uint8)
        status = np.zeros((n, 10), np.uint32)
        msg = np.zeros(n, np.uint32)
        term = np.zeros(n, np.uint8)
        self.L.orc_batch_get_obs(self.ptrs, n, screen.ctypes.data, hist.ctypes.data, status.ctypes.data,
                                 msg.ctypes.data, term.ctypes.data)
        return dict(screen=screen, history=hist, status=status, message=msg, done=term)

    def hashes(self):
        out = np.zeros(self.n, np.uint64)
        self.L.orc_batch_hash(self.ptrs, self.n, out.ctypes.data)
        return out


KEYS11 = np.frombuffer(b".hjklnbuy>s", np.uint8)


def synthetic_actions(t, env_ids):
    """SURVEY §8d action stream: a[i,t] = splitmix64(0x9E3779B97F4A7C15*(t+1) ^ i) % 11 -> ASCII key."""
    with np.errstate(over="ignore"):
        x = (np.uint64(0x9E3779B97F4A7C15) * np.uint64(t + 1)) ^ np.asarray(env_ids, np.uint64)
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return KEYS11[(x % np.uint64(11)).astype(np.int64)]
