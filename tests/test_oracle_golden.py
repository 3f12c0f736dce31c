"""The CPU oracle against every live known answer the reference holds for the path
(SURVEY.md §8c-3). No GPU needed."""
import numpy as np


def _reward_run(oracle, cfg, key_strs, stair_reward):
    """RogueEnv.step(str) + StairRewardEnv / FirstFloorEnv (python/rogue_gym/envs/wrappers.py:12-44)."""
    env = oracle.OracleEnv(cfg)
    rewards, level = [], 1
    for keys in key_strs:
        gold_before = int(env.obs()["status"][1])
        env.react_str(keys)
        st = env.obs()["status"]
        r = int(st[1]) - gold_before
        if int(st[0]) > level:
            level = int(st[0])
            r += stair_reward
        rewards.append(r)
    return env, rewards


def test_seed1_floor_cell_for_cell(oracle, fixtures):
    fx = fixtures["seed1_dungeon_clear"]
    env = oracle.OracleEnv(fx["config"])
    assert env.dungeon() == fx["screen"]


def test_first_floor_env(oracle, fixtures):
    fx = fixtures["first_floor"]
    env, rewards = _reward_run(oracle, fx["config"], [fx["keys"]], fx["stair_reward"])
    assert rewards == [fx["expect_reward"]]
    assert int(env.obs()["status"][0]) == 2  # FirstFloorEnv: done when level 2 is reached
    img = env.encode(1, fx["image_status_flag"], False)
    assert list(img.shape) == fx["expect_image_shape"]


def test_stair_reward_env(oracle, fixtures):
    fx = fixtures["stair_reward"]
    env, rewards = _reward_run(oracle, fx["config"], fx["keys"], fx["stair_reward"])
    assert rewards == fx["expect_rewards"]
    img = env.encode(1, fx["image_status_flag"], True)
    assert list(img.shape) == fx["expect_image_shape"]
    assert img[17][0][0] == fx["expect_img_17_0_0"]
    assert img[18][0][0] == fx["expect_img_18_0_0"]
    st = env.obs()["status"]
    assert [int(st[i]) for i in (0, 2, 3, 4, 5, 6, 7, 8, 9)] == fx["expect_full_status_vec"]


def test_seed1_screens_with_monsters(oracle, fixtures):
    """python/tests/data.py SEED1_DUNGEON2 / SEED1_DUNGEON3: the reference's known answers with the default monster
    table - the screens after 'kHhhKK' and after 'llljln' on seed 1, a snake ('S') beside the player."""
    fx = fixtures["seed1_with_monsters"]
    for case in fx["cases"]:
        e = oracle.OracleEnv(fx["config"])
        e.react_str(case["keys"])
        assert e.dungeon() == case["screen"], case["keys"]
        assert any("S" in row for row in case["screen"])
        mons, _ = e.entities()
        assert len(mons) >= 1


def test_move_enemy_tie_break(oracle, fixtures):
    fx = fixtures["move_enemy"]
    env = oracle.OracleEnv(fx["config"])
    kind, nxt = env.test_move_enemy(tuple(fx["from"]), tuple(fx["to"]))
    assert kind == 1 and list(nxt) == fx["expect"]


def test_noaction_and_max_steps(oracle):
    """python/tests/test_rogue_env.py:28-44: '.' changes nothing; done at max_steps."""
    env = oracle.OracleEnv({"seed": 1}, max_steps=5)
    before = env.obs()
    for i in range(5):
        assert not env.obs()["is_terminal"]
        env.react(".")
    after = env.obs()
    assert after["is_terminal"]
    assert np.array_equal(before["screen"], after["screen"])


def test_same_seed_same_game_and_keys_diverge(oracle):
    """python/src/thread_impls.rs:137-174."""
    envs = [oracle.OracleEnv({"seed": 3}) for _ in range(4)]
    assert len({e.state_hash() for e in envs}) == 1
    for e, k in zip(envs, "hjkl"):
        e.react(k)
    assert len({e.state_hash() for e in envs}) > 1


def test_auto_reset_returns_fresh_state_flagged_terminal(oracle):
    env = oracle.OracleEnv({"seed": 2}, max_steps=3)
    first = env.obs()["screen"].copy()
    for _ in range(2):
        env.step_auto("h")
        assert not env.obs()["is_terminal"]
    env.step_auto("h")
    o = env.obs()
    assert o["is_terminal"] and np.array_equal(o["screen"], first) and env.scalars().steps == 0


def test_errors(oracle):
    import pytest
    env = oracle.OracleEnv({"seed": 2})
    with pytest.raises(oracle.OracleError) as ei:
        env.react("x")
    assert ei.value.code == 1
    assert env.scalars().steps == 0


def test_batch_rollout_is_deterministic(oracle):
    ids = np.arange(64)
    digests = []
    for _ in range(2):
        b = oracle.OracleBatch({}, 64, seeds=list(range(1, 65)), threads=4)
        b.reset()
        for t in range(150):
            b.step(oracle.synthetic_actions(t, ids))
        digests.append(b.hashes())
    assert np.array_equal(digests[0], digests[1])


# ---------------------------------------------------------------- the reference's own unit-test known answers
def test_fenwick_set_known_answers(oracle):
    """core/src/fenwick.rs:262-305 (tests `nth`, `invalid_value`, `from_range`) restated on the oracle's set."""
    cap = 1_000_000
    odd = [i for i in range(1000) if i % 2 == 1]
    assert oracle.ordset_nth(cap, odd, 9)[0] == 19
    assert oracle.ordset_nth(cap, odd, 499)[0] == 999
    v, n, _ = oracle.ordset_nth(cap, odd, 500)
    assert v is None and n == 500
    # out-of-range members are refused by insert and by remove
    bad = list(range(cap, cap + 10))
    _, n, res = oracle.ordset_nth(cap, bad + bad, 0, ops=[1] * 10 + [2] * 10)
    assert n == 0 and not res.any()
    # a second insert of the same member is refused, a remove of a present one succeeds
    _, n, res = oracle.ordset_nth(100, [5, 5, 7, 5, 5], 0, ops=[1, 1, 1, 2, 2])
    assert res.tolist() == [True, False, True, True, False] and n == 1
    L = oracle.lib()
    for i in range(0, 1000, 7):
        assert bool(L.orc_test_ordset_from_range_contains(40, 500, i)) == (40 <= i < 500)


def test_inclusive_edges_known_answer(oracle):
    """core/src/dungeon/rogue/passages.rs:272-296 `test_inclusive_edges`: rect x 5..10, y 6..9."""
    UP, DOWN, LEFT, RIGHT = 0, 1, 2, 3
    assert oracle.edges(5, 10, 6, 9, DOWN, True) == [(x, 8) for x in range(6, 9)]
    assert oracle.edges(5, 10, 6, 9, UP, True) == [(x, 6) for x in range(6, 9)]
    assert oracle.edges(5, 10, 6, 9, LEFT, True) == [(5, y) for y in range(7, 8)]
    assert oracle.edges(5, 10, 6, 9, RIGHT, True) == [(9, y) for y in range(7, 8)]
    # the non-inclusive form (maze exits, passages.rs:160-176) keeps the corners
    assert oracle.edges(5, 10, 6, 9, DOWN, False) == [(x, 8) for x in range(5, 10)]
    assert oracle.edges(5, 10, 6, 9, LEFT, False) == [(5, y) for y in range(6, 9)]


def _ddqn_keys(fixtures, n):
    import json

    from rogue_gym_python._rogue_gym import keys_from_history
    acts = [a for a, k in fixtures["recorded_episode"]["actions_rle"] for _ in range(k)][:n]
    return keys_from_history(json.dumps(acts))


def test_reference_recordings_frame_by_frame(oracle, fixtures, gif_frames):
    """data/gif/*.gif are renderings of the reference's own runs (act2gif), decoded glyph by glyph into
    tests/golden/reference_gif_frames.json. 28 frames of the shipped DDQN episode's first 30 actions and 48 frames of a
    second agent on the same 32x16 game: two gold pick-ups with their amounts, a walk through a passage into a second
    room, the descent, the level-2 floor, three searches of which the third finds the hidden door (the search draws),
    more gold and passages on level 2 - every frame's dungeon rows, status line and messages."""
    from helpers import check_against_recording
    cfg = gif_frames["config"]
    for name, keys in (("ddqn_small", _ddqn_keys(fixtures, gif_frames["ddqn_small"]["n_actions"])),
                       ("ppo_cog19", gif_frames["ppo_cog19"]["keys"].encode())):
        frames = gif_frames[name]["frames"]
        e = oracle.OracleEnv(cfg, max_steps=2000)
        n = check_against_recording(e.react, e.dungeon, lambda: e.obs()["status"], lambda: e.obs()["message"], keys, frames)
        assert n == len(frames), (name, n, len(frames))
        assert min(gif_frames[name]["worst_glyph_score"]) > 0.6
    last = gif_frames["ppo_cog19"]["frames"][-1]
    assert last[0].startswith("Hmm... there seems to be no down") and last[-1].startswith("Level:  2 Gold:    12 Hp: 12(12)")


def test_dice_range(oracle):
    """core/src/character/mod.rs:277-285 `test_dice`: 2d4 stays in 2..=8 (here also: reaches both ends, and the mean of
    a fair 2d4 is 5)."""
    for seed in (1, 2, 12345, 2**63 + 7):
        r = oracle.dice(seed, 2, 4, 100)
        assert r.min() >= 2 and r.max() <= 8
    r = oracle.dice(9, 2, 4, 20000)
    assert r.min() == 2 and r.max() == 8 and abs(r.mean() - 5.0) < 0.05
    assert (oracle.dice(3, 0, 4, 10) == 0).all()  # no dice, no damage


def test_select_cell_runs_dry(oracle):
    """core/src/dungeon/rogue/floor.rs:490-505 `select_cell`: on a level-10 floor of the default config, select_cell
    + set_obj hands out more than 15 cells before it returns None, and set_obj never refuses a selected cell."""
    for seed in range(1, 41):
        n = oracle.select_cells({}, 10, seed)
        assert 15 < n < 1000, (seed, n)


def test_gif_fixture_regenerates_from_the_reference(gif_frames):
    """Where the reference tree and PIL are present (this container, not the GPU box): decoding data/gif/*.gif again
    gives the committed frames, the 16 px and 24 px renderings agree, and the inferred key sequence comes out the same."""
    import importlib.util
    import os

    import pytest
    if not os.path.isdir("/root/reference/data/gif") or importlib.util.find_spec("PIL") is None:
        pytest.skip("needs /root/reference and PIL")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_gif_golden.py")
    spec = importlib.util.spec_from_file_location("make_gif_golden", here)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    gif = os.path.join(m.REF, "data/gif")
    f16, _ = m.decode(os.path.join(gif, "ddqn-small-16.gif"), 16)
    f24, _ = m.decode(os.path.join(gif, "ddqn-small-24.gif"), 24)
    assert f16 == f24 == gif_frames["ddqn_small"]["frames"]
    ppo, _ = m.decode(os.path.join(gif, "pporesnet-cog19-10seed.gif"), 16)
    assert ppo == gif_frames["ppo_cog19"]["frames"]
    assert m.infer_actions(gif_frames["config"], ppo) == gif_frames["ppo_cog19"]["keys"]
