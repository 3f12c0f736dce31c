"""ctypes declarations of include/rogue_b200.h (the C ABI of librogue_b200.so).

Nothing here computes: every call lands in the CUDA library. If the library is missing the
import fails loudly; if no GPU is present every compute entry point returns RG_ERR_CUDA, which
`check()` raises as RuntimeError. There is no CPU fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ROGUE_B200_LIB") or os.path.normpath(os.path.join(HERE, "..", "..", "librogue_b200.so"))

MAX_ENEMY_KINDS, MAX_DICE, MAX_EXPS, MAX_INIT_DRAWS, MAX_ROOMS, DIST_CACHE = 32, 4, 32, 8, 16, 9
(RG_OK, RG_ERR_INVALID_INPUT, RG_ERR_IGNORED_INPUT, RG_ERR_PANIC, RG_ERR_SETTING, RG_ERR_PARSE, RG_ERR_CUDA,
 RG_ERR_ARG) = range(8)


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
        ("has_seed", C.c_int32), ("seed_lo", C.c_uint64), ("seed_hi", C.c_uint64),
        ("has_seed_range", C.c_int32), ("seed_range_lo", C.c_uint64), ("seed_range_hi", C.c_uint64),
    ]


class Views(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int64), ("width", C.c_int32), ("height", C.c_int32),
        ("cell_stride", C.c_int32), ("hist_stride", C.c_int32),
        ("screen", C.c_void_p), ("history_bits", C.c_void_p), ("status", C.c_void_p), ("reward", C.c_void_p),
        ("done", C.c_void_p), ("message", C.c_void_p), ("error", C.c_void_p),
    ]


class HostObs(C.Structure):
    _fields_ = [
        ("screen", C.c_void_p), ("history", C.c_void_p), ("status", C.c_void_p), ("reward", C.c_void_p),
        ("done", C.c_void_p), ("message", C.c_void_p), ("error", C.c_void_p),
    ]


class DumpScalars(C.Structure):
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


class Dump(C.Structure):
    _fields_ = [
        ("s", DumpScalars), ("surface", C.c_void_p), ("attr", C.c_void_p), ("monsters", C.c_void_p),
        ("items", C.c_void_p), ("cache_xy", C.c_void_p), ("cache_maps", C.c_void_p), ("rooms", C.c_void_p),
    ]


# every symbol include/rogue_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _u32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32
SYMBOLS = {
    "rg_parse_config": (_i, [C.c_char_p, C.POINTER(Params), C.c_char_p, C.c_size_t]),
    "rg_validate_params": (_i, [C.POINTER(Params), C.c_char_p, C.c_size_t]),
    "rg_create": (_i, [C.POINTER(C.c_char_p), _i64, _i64, _i64, _i, C.POINTER(_vp)]),
    "rg_create_from_params": (_i, [C.POINTER(Params), _i64, _i64, _i, C.POINTER(_vp)]),
    "rg_destroy": (None, [_vp]),
    "rg_last_error": (C.c_char_p, [_vp]),
    "rg_version": (C.c_char_p, []),
    "rg_seed": (_i, [_vp, _vp, _vp]),
    "rg_seed_first": (_i, [_vp, _vp, _vp, _i64]),
    "rg_reset": (_i, [_vp]),
    "rg_step": (_i, [_vp, _vp, _i]),
    "rg_step_host": (_i, [_vp, _vp, _i, C.POINTER(HostObs)]),
    "rg_sync": (_i, [_vp]),
    "rg_quiesce": (_i, [_vp]),
    "rg_stats": (_i, [_vp, _vp]),
    "rg_trace": (_i, [_vp, _vp, _vp]),
    "rg_views_get": (_i, [_vp, C.POINTER(Views)]),
    "rg_fetch": (_i, [_vp, C.POINTER(HostObs)]),
    "rg_fetch_terminal": (_i, [_vp, _vp]),
    "rg_set_panic_policy": (_i, [_vp, _i]),
    "rg_step_train": (_i, [_vp, _vp, _i, _i, _u32, _i, _vp, _vp, C.c_float]),
    "rg_train_reset": (_i, [_vp]),
    "rg_mirror_get": (_i, [_vp, C.POINTER(HostObs), C.POINTER(_vp)]),
    "rg_mirror_sync": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "rg_step_mirror": (_i, [_vp, _vp, _i, C.POINTER(C.c_uint64)]),
    "rg_stream": (_vp, [_vp]),
    "rg_launch_count": (_i64, [_vp]),
    "rg_encode": (_i, [_vp, _i, _u32, _i, _vp, C.POINTER(_i)]),
    "rg_encode_channels": (_i, [_vp, _i, _u32, _i]),
    "rg_encode_compact": (_i, [_vp, _vp, _vp, _vp]),
    "rg_encode_states": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _u32, _i, _vp, C.POINTER(_i)]),
    "rg_dump_env": (_i, [_vp, _i64, C.POINTER(Dump)]),
    "rg_state_hash": (_i, [_vp, _vp]),
    "rg_export_floors": (_i, [_vp, _vp, _vp, _vp]),
    "rg_test_move_enemy": (_i, [_vp, _i64, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
}

_lib = None


def lib():
    """Loads librogue_b200.so (built in-tree by `make -C rogue-gym_b200` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "librogue_b200.so not found at %s: build it with `make -C rogue-gym_b200` "
                "(there is no pure-Python or CPU implementation to fall back to)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error(handle=None):
    msg = lib().rg_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, handle=None, tolerate_env_errors=False):
    """Errors surface the way the PyO3 layer raises them: RuntimeError (python/src/lib.rs:20-26).
    Batched callers that read the per-env `error` array pass tolerate_env_errors=True."""
    if tolerate_env_errors and rc in (RG_ERR_INVALID_INPUT, RG_ERR_IGNORED_INPUT, RG_ERR_PANIC):
        return
    if rc != RG_OK:
        err = RuntimeError(last_error(handle) or ("rogue-gym_b200 error %d" % rc))
        err.code = rc
        raise err
