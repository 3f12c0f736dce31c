"""Host-side checks of the C-ABI library that need no GPU: it loads, exports every symbol
include/rogue_b200.h declares, parses the GameConfig schema exactly like the oracle's
independent parser, validates sizes like the reference, and fails loudly without a device."""
import ctypes as C
import json
import os
import re
import subprocess

import pytest

from helpers import CONFIGS, ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rogue_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(cabi):
    names = _declared_symbols()
    assert len(names) >= 20
    L = C.CDLL(cabi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "librogue_b200.so does not export %s" % n
    assert set(names) == set(cabi.SYMBOLS), "ctypes table and header disagree"


def test_library_holds_sm100a_code_only(cabi):
    out = subprocess.run(["cuobjdump", "-lelf", cabi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "rogue_b200.h"\nint main(void){ rg_params p; (void)p; return (int)sizeof(rg_views) * 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def _fields(struct):
    out = {}
    for name, _ in struct._fields_:
        v = getattr(struct, name)
        if hasattr(v, "_fields_"):
            v = _fields(v)
        elif hasattr(v, "__len__"):
            v = [(_fields(x) if hasattr(x, "_fields_") else x) for x in v]
        out[name] = v
    return out


EXTRA = {
    "reference_default_json": None,  # filled from tests/golden/config_default.json
    "custom_monsters": {"enemies": {"enemies": [
        1, {"attack": [{"times": 2, "max": 6}], "attr": 513, "defense": 4, "exp": 9, "gold": 0, "level": 2,
            "name": "thing", "tile": 84, "rarelity": 1}, 10], "appear_rate_gold": 100, "appear_rate_nogold": 60}},
    "custom_player": {"player": {"exps": [5, 10, 4294967295], "hunger_time": 200, "init_hp": 30, "max_items": 3,
                                 "init_items": [{"Weapon": {"name": "dagger", "num_plus": 0, "hit_plus": 2, "dam_plus": 3}},
                                                {"Armor": {"name": "plate mail", "def_plus": -1}},
                                                {"Noinit": {"kind": "Gold", "how_many": 7, "attr": 4}}]}},
    "no_gold_slot": {"player": {"max_items": 1, "init_items": [
        {"Noinit": {"kind": {"Food": "Ration"}, "how_many": 1, "attr": 4}}]}},
    "custom_items": {"item": {"gold": {"rate_inv": 1, "base": 10, "per_level": 1, "minimum": 5},
                              "weapon": {"weapons": [0, 3, 2]}, "armor": {"armors": [1]}}},
    "big_seed": {"seed": 340282366920938463463374607431768211455},
    "seed_range": {"seed_range": [10, 20]},
}


@pytest.mark.parametrize("name", sorted(CONFIGS) + sorted(EXTRA))
def test_parser_matches_oracle_parser(cabi, oracle, name):
    if name == "reference_default_json":
        with open(os.path.join(ROOT, "tests", "golden", "config_default.json")) as f:
            cfg = json.load(f)
    else:
        cfg = CONFIGS.get(name, EXTRA.get(name))
    want, seed = oracle.params_from_config(cfg)
    got = cabi.Params()
    err = C.create_string_buffer(256)
    rc = cabi.lib().rg_parse_config(json.dumps(cfg).encode(), C.byref(got), err, 256)
    assert rc == 0, err.value
    w, g = _fields(want), _fields(got)
    for k in w:
        assert g[k] == w[k], k
    assert bool(g["has_seed"]) == (seed is not None)
    if seed is not None:
        assert g["seed_lo"] | (g["seed_hi"] << 64) == seed
    if "seed_range" in cfg and cfg["seed_range"]:
        assert (g["has_seed_range"], g["seed_range_lo"], g["seed_range_hi"]) == (1, *cfg["seed_range"])
    assert cabi.lib().rg_validate_params(C.byref(got), err, 256) == 0, err.value


@pytest.mark.parametrize("text,needle", [
    ("{", "Failed to parse config"),
    ('{"width": "80"}', "Failed to parse config"),
    ('{"dungeon": {"room_num_x": 2}}', "style"),
    ('{"dungeon": {"style": "nethack"}}', "nethack"),
    ('{"player": {"init_items": [{"Weapon": {"name": "spoon", "num_plus": 0, "hit_plus": 0, "dam_plus": 0}}]}}', "spoon"),
    ('{"enemies": {"enemies": [26]}}', "out of range"),
    ('{"width": 80} x', "trailing"),
])
def test_parse_errors(cabi, text, needle):
    p = cabi.Params()
    err = C.create_string_buffer(512)
    assert cabi.lib().rg_parse_config(text.encode(), C.byref(p), err, 512) == cabi.RG_ERR_PARSE
    assert needle in err.value.decode()


@pytest.mark.parametrize("cfg,needle", [
    ({"width": 31}, "screen width is too narrow"),   # core/src/lib.rs:167-184
    ({"width": 161}, "screen width is too wide"),
    ({"height": 15}, "screen height is too narrow"),
    ({"height": 49}, "screen height is too wide"),
    ({"width": 32, "height": 16}, "room grid does not fit"),  # 3x3 rooms at 32x16: range(4..4) panics (rooms.rs:256-259)
    ({"dungeon": {"style": "rogue", "max_extra_edges": 0}}, "is 0"),
])
def test_validate_like_reference(cabi, cfg, needle):
    p = cabi.Params()
    err = C.create_string_buffer(512)
    assert cabi.lib().rg_parse_config(json.dumps(cfg).encode(), C.byref(p), err, 512) == 0
    assert cabi.lib().rg_validate_params(C.byref(p), err, 512) == cabi.RG_ERR_SETTING
    assert needle in err.value.decode()


def test_no_device_fails_loudly(cabi):
    """The product has no CPU path: without a GPU, creating a batch is an error, never a fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rogue_gym_python import _rogue_gym
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _rogue_gym.GameState(10, "{}")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _rogue_gym.ParallelGameState(10, ["{}"] * 2)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, include or link it."""
    pkg = os.path.join(ROOT, "rogue-gym_b200")
    for dp, _, fs in os.walk(pkg):
        if "/build" in dp:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle_py" not in text and "liboracle" not in text and "oracle.h" not in text, os.path.join(dp, f)
    ldd = subprocess.run(["ldd", os.path.join(pkg, "librogue_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_dump_config_round_trip():
    """python/tests/test_ff_env.py:22: get_config() == CONFIG."""
    from rogue_gym_python._config import dump_config
    cfg = {"seed": 1, "hide_dungeon": False, "enemies": {"enemies": []}}
    assert json.loads(dump_config(json.dumps(cfg))) == cfg
    assert json.loads(dump_config("{}")) == {"hide_dungeon": True}
    assert json.loads(dump_config("{}", 9)) == {"seed": 9, "hide_dungeon": True}
    d = json.loads(dump_config(json.dumps({"dungeon": {"style": "rogue", "room_num_x": 2}})))
    assert d["dungeon"]["room_num_x"] == 2 and d["dungeon"]["room_num_y"] == 3 and d["dungeon"]["amulet_level"] == 25


def test_builtin_defaults_equal_the_shipped_default_config(cabi):
    """data/config-default.json is what GameConfig::default() serialises to (core/src/lib.rs:438-445 `print_default`):
    every built-in table of the product's parser - the 26 monsters, weapons, armors, item rates, player and dungeon
    parameters - must equal it field for field. One known difference: the file's level table ends in 0 where the
    source has u32::max_value() (character/player.rs:336; the file predates it)."""
    def parse(cfg):
        p = cabi.Params()
        err = C.create_string_buffer(256)
        assert cabi.lib().rg_parse_config(json.dumps(cfg).encode(), C.byref(p), err, 256) == 0, err.value
        return _fields(p)

    with open(os.path.join(ROOT, "tests", "golden", "config_default.json")) as f:
        shipped = parse(json.load(f))
    builtin = parse({})
    assert len(builtin) >= 40
    for k in builtin:
        if k == "exps":
            n = [i for i, v in enumerate(builtin[k]) if v == 4294967295]
            assert len(n) == 1 and shipped[k][n[0]] == 0
            assert builtin[k][:n[0]] == shipped[k][:n[0]] and builtin[k][n[0] + 1:] == shipped[k][n[0] + 1:]
        else:
            assert builtin[k] == shipped[k], k
