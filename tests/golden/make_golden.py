"""Writes tests/golden/reference_fixtures.json from the reference's own test data.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
It imports the reference's python/tests/data.py (pure data, no extension module needed) and
records, next to each vector, the reference test that uses it and the expectation that test
asserts. The Rust extension itself cannot be built here (no rustc/cargo), so these live
fixtures are the reference's only executable-independent known answers.
"""
import json
import os
import sys

REF = "/root/reference/python/tests"
sys.path.insert(0, REF)
import data  # noqa: E402

out = {
    "source": "kngwyu/rogue-gym @ c78608b python/tests/data.py",
    "seed1_dungeon_clear": {
        "cite": "python/tests/data.py:83-108, used by python/tests/test_ff_env.py:14-16",
        "config": {"seed": 1, "hide_dungeon": False, "enemies": {"enemies": []}},
        "screen": data.SEED1_DUNGEON_CLEAR,
    },
    "first_floor": {
        "cite": "python/tests/test_ff_env.py:5-22",
        "config": {"seed": 1, "hide_dungeon": False, "enemies": {"enemies": []}},
        "keys": data.CMD_STR2,
        "stair_reward": 100.0,
        "expect_reward": 102,
        "expect_done": True,
        "image_status_flag": 1,
        "expect_image_shape": [18, 24, 80],
    },
    "stair_reward": {
        "cite": "python/tests/test_st_env.py:11-37",
        "config": {"width": 32, "height": 16, "seed": 5, "hide_dungeon": False,
                   "dungeon": {"style": "rogue", "room_num_x": 2, "room_num_y": 2},
                   "enemies": {"enemies": []}},
        "keys": [data.CMD_STR3, data.CMD_STR4],
        "stair_reward": 100.0,
        "expect_rewards": [104.0, 100.0],
        "image_status_flag": 1 | 2 | 128,
        "expect_image_shape": [21, 16, 32],
        "expect_img_17_0_0": 3.0,
        "expect_img_18_0_0": 12.0,
        "expect_full_status_vec": [3, 12, 12, 16, 16, 0, 1, 0, 0],
    },
    "move_enemy": {
        "cite": "core/src/dungeon/rogue/mod.rs:566-578 (test_move_enemy)",
        "config": {"width": 32, "height": 16, "seed": 5,
                   "dungeon": {"style": "rogue", "room_num_x": 2, "room_num_y": 2}},
        "from": [9, 9], "to": [28, 4], "expect": [10, 9],
    },
    # The reference's only known answers WITH MONSTERS ENABLED (default config, seed 1): the screens after CMD_STR and
    # after CMD_STR5. SURVEY.md 8c-3 filed them under "stale" together with SEED1_DUNGEON; they are not - the oracle and
    # the GPU reproduce both screens cell for cell, including the monster 'S' next to the player. They pin: the enemy
    # stream's gen_enemy draws on level 1 (which room gets which kind, where), MoveUntil through a dark room, the
    # 9-neighbour visibility of a dark room, monster drawing (Dungeon::draw_enemy), the item stream's init draws.
    "seed1_with_monsters": {
        "cite": "python/tests/data.py:25-81 (CMD_STR, CMD_STR5, SEED1_DUNGEON2, SEED1_DUNGEON3), used by "
                "python/tests/test_rogue_env.py:21-25 and test_parallel.py:27-40",
        "config": {"seed": 1},
        "cases": [{"keys": data.CMD_STR, "screen": data.SEED1_DUNGEON2},
                  {"keys": data.CMD_STR5, "screen": data.SEED1_DUNGEON3}],
    },
    "stale": {
        "note": "SEED1_DUNGEON (21 rows; the API returns 24, and the start room sits one row lower than in the live golden) "
                "is stale in the reference itself (SURVEY 8c-3): test_rogue_env.py::test_screen and the test_parallel.py "
                "cases that first compare with it cannot pass against the reference's own code",
        "seed1_dungeon_rows": len(data.SEED1_DUNGEON),
    },
}
# A recorded episode the reference ships (RunTime::saved_inputs_as_json format, core/src/lib.rs:357-375):
# kept as its config plus the action list in the reference's own JSON shape, run-length encoded.
_learned = "/root/reference/data/learned/ddqn-minidungeon"
with open(os.path.join(_learned, "best-actions.json")) as f:
    _acts = json.load(f)
_rle = []
for a in _acts:
    if _rle and _rle[-1][0] == a:
        _rle[-1][1] += 1
    else:
        _rle.append([a, 1])
with open(os.path.join(_learned, "config.json")) as f:
    _cfg = json.load(f)
out["recorded_episode"] = {
    "cite": "data/learned/ddqn-minidungeon/{best-actions,config}.json; format core/src/lib.rs:357-375, input.rs:24-60",
    "config": _cfg, "n_actions": len(_acts), "actions_rle": _rle,
}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.json")
with open(dst, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", dst)
