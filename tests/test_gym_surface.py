"""The gym layer: `rogue_gym.envs` on top of the B200 module.

The GPU tests restate the reference's own live pytest cases (python/tests/test_rogue_env.py,
test_ff_env.py, test_st_env.py, test_parallel.py) with the expectations recorded in
tests/golden/reference_fixtures.json, then check the device-resident env against the list API.
The reference's SEED1_DUNGEON/2/3 vectors are stale (21 rows), so where the
reference compares against them these tests compare against the oracle instead.
"""
import json

import numpy as np
import pytest


# --------------------------------------------------------------------------- CPU: host logic only
def test_gym_api_resolves_and_spaces_compare_by_value():
    from rogue_gym import _gymapi
    sp = _gymapi.spaces
    assert sp.discrete.Discrete(11) == sp.discrete.Discrete(11)
    assert sp.discrete.Discrete(11) != sp.discrete.Discrete(10)
    a = sp.box.Box(low=0, high=1, shape=(26, 24, 80), dtype=np.float32)
    assert a == sp.box.Box(low=0, high=1, shape=(26, 24, 80), dtype=np.float32)
    assert a != sp.box.Box(low=0, high=1, shape=(25, 24, 80), dtype=np.float32)
    import gym  # the reference's tests import it by this name (test_rogue_env.py:3-4)
    assert gym.spaces.discrete.Discrete(3) == sp.discrete.Discrete(3)


def test_gym_shim_wrapper_semantics_and_optional_adapter():
    from rogue_gym import _gymapi
    import rogue_gym  # the package imports without `rainy`

    class Dummy(_gymapi.Env):
        action_space = _gymapi.spaces.discrete.Discrete(3)
        observation_space = _gymapi.spaces.box.Box(low=0, high=1, shape=(2, 2), dtype=np.float32)
        extra = 41

        def step(self, a):
            return a, 1.0, False, {}

        def reset(self):
            return 0

    w = _gymapi.Wrapper(Dummy())
    assert w.unwrapped is w.env and w.action_space == Dummy.action_space and w.extra == 41
    assert w.step(2) == (2, 1.0, False, {}) and w.reset() == 0
    if _gymapi.IS_SHIM:
        assert w.action_space.contains(2) and not w.action_space.contains(3)
        assert w.observation_space.sample().shape == (2, 2)
    assert rogue_gym.__version__.startswith("0.0.2")


def test_status_flag_and_image_setting_dims():
    from rogue_gym.envs import DungeonType, ImageSetting, StatusFlag
    assert StatusFlag.FULL.value == 0b111111111 and StatusFlag.FULL.count_one() == 9
    assert (StatusFlag.DUNGEON_LEVEL | StatusFlag.HP_CURRENT | StatusFlag.EXP).count_one() == 3
    assert StatusFlag.EMPTY.count_one() == 0
    assert ImageSetting().dim(43) == 52  # reference default: 43 symbol planes + 9 status planes
    assert ImageSetting(DungeonType.GRAY, StatusFlag.EMPTY, True).dim(43) == 2
    assert ImageSetting(DungeonType.SYMBOL, StatusFlag.DUNGEON_LEVEL, False).dim(17) == 18
    assert ImageSetting().encoder_args() == (1, 0x1FF, 0)
    with pytest.raises(TypeError):
        ImageSetting().expand("not a state")
    with pytest.raises(TypeError):
        StatusFlag.FULL.status_vec(None)


def test_action_tables_match_the_keymap():
    from rogue_gym.envs import ParallelRogueEnv, RogueEnv
    assert "".join(RogueEnv.ACTIONS) == ".hjklnbuy>s" and RogueEnv.ACTION_LEN == 11
    assert set(RogueEnv.ACTION_MEANINGS) == set(RogueEnv.ACTIONS)
    assert len(RogueEnv.SYMBOLS) == 43 and RogueEnv.SYMBOLS[1] == "@" and RogueEnv.SYMBOLS[17] == "A"
    assert ParallelRogueEnv.ACTIONS is RogueEnv.ACTIONS


def test_env_needs_the_gpu_and_says_so(cabi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from rogue_gym.envs import RogueEnv
    with pytest.raises(RuntimeError):
        RogueEnv(seed=1)


def _recorded(fixtures):
    fx = fixtures["recorded_episode"]
    acts = [a for a, n in fx["actions_rle"] for _ in range(n)]
    assert len(acts) == fx["n_actions"]
    return fx["config"], acts


def test_action_history_round_trip(fixtures):
    from rogue_gym_python._rogue_gym import _input_code, keys_from_history
    _, acts = _recorded(fixtures)
    keys = keys_from_history(json.dumps(acts))
    assert len(keys) == 1000 and set(keys) <= set(b"hjklyubn.s>")
    assert [_input_code(k) for k in keys] == acts  # the writer side reproduces the reference's file
    assert keys_from_history('[{"Act": {"MoveUntil": "Left"}}, {"Act": "NoOp"}]') == b"H."
    with pytest.raises(ValueError):
        keys_from_history('[{"Sys": "Quit"}]')


# --------------------------------------------------------------------------- GPU: reference cases
@pytest.fixture()
def envs(gpu):
    import rogue_gym.envs as E
    return E


@pytest.mark.gpu
def test_screen_and_kwargs(envs):  # test_rogue_env.py:16-22,43-45
    env = envs.RogueEnv(seed=1)
    assert env.screen_size() == (24, 80)
    rows = env.get_dungeon()
    assert len(rows) == 24 and all(len(r) == 80 for r in rows)
    assert sum(r.count("@") for r in rows) == 1
    assert envs.RogueEnv(seed=1, width=48, height=24).screen_size() == (24, 48)
    assert envs.RogueEnv().screen_size() == (24, 80)  # keyword settings do not leak into later envs


@pytest.mark.gpu
def test_noaction_and_max_steps(envs, fixtures):  # test_rogue_env.py:31-41
    env = envs.RogueEnv(seed=1)
    before = env.result
    after, reward, done, info = env.step(".")
    assert after.dungeon == before.dungeon and after.status == before.status
    assert reward == 0 and not done and info == {}
    env = envs.RogueEnv(seed=1, max_steps=5)
    _, _, done, _ = env.step(fixtures["seed1_with_monsters"]["cases"][0]["keys"])
    assert done
    with pytest.raises(ValueError):
        env.step(11)


@pytest.mark.gpu
def test_action_string_against_oracle(envs, fixtures, oracle):  # test_rogue_env.py:21-25 (its golden, SEED1_DUNGEON2, is live)
    env = envs.RogueEnv(seed=1)
    keys = fixtures["seed1_with_monsters"]["cases"][0]["keys"]
    res, *_ = env.step(keys)
    o = oracle.OracleBatch({"seed": 1}, 1)
    o.reset()
    for k in keys:
        o.step(np.array([ord(k)], np.uint8), False)
    W = 80
    want = [bytes(o.obs()["screen"][0][i:i + W]).decode("latin-1") for i in range(0, 24 * W, W)]
    assert res.dungeon == want == fixtures["seed1_with_monsters"]["cases"][0]["screen"]


@pytest.mark.gpu
def test_images_and_spaces(envs):  # test_rogue_env.py:48-68
    cfg = {"seed": 1, "enemies": {"enemies": []}}
    env = envs.RogueEnv(config_dict=cfg)
    state, *_ = env.step("H")
    empty = envs.StatusFlag.EMPTY
    img = empty.symbol_image_with_hist(state)
    assert img.shape == (18, 24, 80) and img.dtype == np.float32
    assert set(np.unique(img[-1])) <= {0.0, 1.0} and img[-1].sum() > 0
    assert empty.gray_image(state).shape == (1, 24, 80)
    assert empty.gray_image_with_hist(state).shape == (2, 24, 80)
    import gym
    from gym import spaces
    assert env.action_space == gym.spaces.discrete.Discrete(env.ACTION_LEN)
    assert env.observation_space == spaces.box.Box(low=0, high=1, shape=(26, 24, 80), dtype=np.float32)
    assert env.get_key_to_action()["h"] == "MOVE_LEFT"


@pytest.mark.gpu
def test_first_floor_env(envs, fixtures):  # test_ff_env.py:5-22
    fx = fixtures["first_floor"]
    setting = envs.ImageSetting(status=envs.StatusFlag(fx["image_status_flag"]))
    env = envs.FirstFloorEnv(envs.RogueEnv(config_dict=fx["config"], image_setting=setting), fx["stair_reward"])
    assert env.unwrapped.get_dungeon() == fixtures["seed1_dungeon_clear"]["screen"]
    state, reward, done, _ = env.step(fx["keys"])
    assert done == fx["expect_done"]
    assert reward == fx["expect_reward"]
    assert env.unwrapped.state_to_image(state).shape == tuple(fx["expect_image_shape"])
    assert env.unwrapped.get_config() == fx["config"]
    assert repr(env) == repr(env.unwrapped.result)


@pytest.mark.gpu
def test_stair_reward_env(envs, fixtures):  # test_st_env.py:11-37
    fx = fixtures["stair_reward"]
    setting = envs.ImageSetting(envs.DungeonType.SYMBOL, envs.StatusFlag(fx["image_status_flag"]), True)
    env = envs.StairRewardEnv(envs.RogueEnv(config_dict=fx["config"], image_setting=setting), fx["stair_reward"])
    state, r1, _, _ = env.step(fx["keys"][0])
    assert r1 == fx["expect_rewards"][0]
    state, r2, _, _ = env.step(fx["keys"][1])
    assert r2 == fx["expect_rewards"][1]
    img = env.unwrapped.state_to_image(state)
    assert img.shape == tuple(fx["expect_image_shape"])
    assert img[17][0][0] == fx["expect_img_17_0_0"] and img[18][0][0] == fx["expect_img_18_0_0"]
    assert envs.StatusFlag.FULL.status_vec(state) == fx["expect_full_status_vec"]
    env.reset()
    assert env.current_level == 1
    with pytest.raises(ValueError):
        from rogue_gym._gymapi import Env
        envs.StairRewardEnv(Env())


@pytest.mark.gpu
def test_save_actions_and_config(envs, tmp_path):
    env = envs.RogueEnv(seed=3)
    env.step("hjL>s.")
    env.save_actions(str(tmp_path / "a.json"))
    env.save_config(str(tmp_path / "c.json"))
    acts = json.load(open(tmp_path / "a.json"))
    assert acts[0] == {"Act": {"Move": "Left"}} and acts[2] == {"Act": {"MoveUntil": "Right"}}
    assert acts[3:] == [{"Act": "DownStair"}, {"Act": "Search"}, {"Act": "NoOp"}]
    assert json.load(open(tmp_path / "c.json")) == {"seed": 3, "hide_dungeon": True}
    with pytest.raises(RuntimeError):
        env.replay()


@pytest.mark.gpu
def test_recorded_episode_replays_like_the_oracle(envs, fixtures, oracle, tmp_path):
    """data/learned/ddqn-minidungeon/best-actions.json (1 000 inputs, reference format) re-simulated on
    the GPU from its own config: same final screen / status as the oracle, and dump_history gives
    the file back."""
    cfg, acts = _recorded(fixtures)
    path = tmp_path / "best-actions.json"
    path.write_text(json.dumps(acts))
    env = envs.RogueEnv(config_dict=cfg, max_steps=2000)
    state, reward, done, _ = env.replay_actions(str(path))
    from rogue_gym_python._rogue_gym import keys_from_history
    o = oracle.OracleEnv(cfg, max_steps=2000)
    for k in keys_from_history(json.dumps(acts)):
        o.react(k)
    assert state.dungeon == o.dungeon()
    assert list(state.status.values()) == [int(v) for v in o.obs()["status"]]
    assert not done and reward == state.gold
    assert json.loads(env.game.dump_history()) == acts
    got = env.get_config()
    assert got["dungeon"] == cfg["dungeon"] and got["enemies"] == cfg["enemies"] and got["seed"] == 5


NUM_WORKERS = 8


@pytest.mark.gpu
def test_parallel_configs_seed_and_cycle(envs, fixtures):  # test_parallel.py:27-62
    cmd, cmd5 = fixtures["seed1_with_monsters"]["cases"][0]["keys"], fixtures["seed1_with_monsters"]["cases"][1]["keys"]
    env = envs.ParallelRogueEnv(config_dicts=[{"seed": 1}] * NUM_WORKERS)
    first = env.states[0].dungeon
    assert all(s.dungeon == first for s in env.states)
    single = [envs.RogueEnv(seed=1), envs.RogueEnv(seed=1)]
    for i in range(len(cmd)):
        env.step("".join((cmd, cmd5)[x % 2][i] for x in range(NUM_WORKERS)))
    single[0].step(cmd)
    single[1].step(cmd5[:len(cmd)])
    for i, s in enumerate(env.states):  # every worker followed its own key string
        assert s.dungeon == single[i % 2].result.dungeon
    env.seed([10] * env.num_workers)
    assert all(s.dungeon != first for s in env.reset())
    assert env.get_configs() == [{"seed": 1, "hide_dungeon": True}] * NUM_WORKERS
    env.close()

    env = envs.ParallelRogueEnv(config_dicts=[{"seed": 1}] * NUM_WORKERS, max_steps=5)
    for i, c in enumerate(cmd):
        states, _, dones, infos = env.step(c * NUM_WORKERS)
        if i % 5 == 4:  # the terminal step returns the fresh game of the next episode
            assert dones == [True] * NUM_WORKERS
            assert all(s.dungeon == first for s in states)
        else:
            assert dones == [False] * NUM_WORKERS
        assert infos == [{}] * NUM_WORKERS
    with pytest.raises(ValueError):
        env.step([0] * (NUM_WORKERS - 1) + [99])


@pytest.mark.gpu
def test_stair_reward_parallel(envs, fixtures):  # test_parallel.py:65-79
    fx = fixtures["stair_reward"]
    env = envs.StairRewardParallel(config_dicts=[fx["config"]] * NUM_WORKERS, max_steps=30)
    for keys in fx["keys"]:
        for c in keys:
            _, rewards, *_ = env.step(c * NUM_WORKERS)
            assert all(r >= 0.0 for r in rewards)
        assert rewards == [50.0] * NUM_WORKERS
    for _ in range(30 - sum(len(k) for k in fx["keys"])):
        _, rewards, dones, _ = env.step([0] * NUM_WORKERS)
        assert all(r >= 0.0 for r in rewards)
    assert dones == [True] * NUM_WORKERS


# --------------------------------------------------------------------------- GPU: device-resident env
def _panic_free_seeds(oracle, cfg, n, steps, max_steps):
    """First n seeds whose rollout under the per-seed action stream never reaches a state where the
    reference panics, so that the list API does not raise mid-test."""
    cand = np.arange(1, 3 * n + 1, dtype=np.uint64)
    ob = oracle.OracleBatch(cfg, len(cand), max_steps=max_steps, seeds=[int(s) for s in cand])
    ob.reset()
    ok = ob.rc == 0
    for t in range(steps):
        ob.step(oracle.synthetic_actions(t, cand), True)
        ok &= ob.rc == 0
    good = cand[ok][:n]
    assert len(good) == n
    return good


@pytest.mark.gpu
@pytest.mark.parametrize("observation", ["compact", "image"])
@pytest.mark.parametrize("setting_args", [(2, 0x1FF, False), (1, 0x03, True), (2, 0, True)])
def test_device_env_matches_list_api(envs, oracle, setting_args, observation):
    """Both observation forms of DeviceRogueEnv against the list API: "image" is the float32 image itself; "compact"
    (the default: symbol ids + status vector [+ visited map], 1.9 KB instead of 401 KB per env) must expand on the
    device - env.expand = SymbolExpand - to exactly that image, every step."""
    import torch
    n, steps, max_steps = 96, 120, 40
    setting = envs.ImageSetting(envs.DungeonType(setting_args[0]), envs.StatusFlag(setting_args[1]), setting_args[2])
    cfg = {}
    seeds = _panic_free_seeds(oracle, cfg, n, steps, max_steps)
    dev = envs.DeviceRogueEnv(cfg, num_envs=n, max_steps=max_steps, image_setting=setting, seeds=seeds, stair_reward=50.0,
                              observation=observation)
    assert envs.DeviceRogueEnv.__init__.__defaults__[-1] == "compact"
    ref = envs.StairRewardParallel(config_dicts=[cfg] * n, max_steps=max_steps, image_setting=setting)
    ref.seed([int(s) for s in seeds])
    states = ref.reset()
    obs = dev.reset()
    if observation == "compact":
        raw = obs
        assert raw.symbols.is_cuda and raw.symbols.dtype == torch.uint8 and raw.symbols.shape == (n, 24, 80)
        assert raw.status.shape == (n, 9) and (raw.history is not None) == setting.includes_hist
        assert int(raw.symbols.max()) <= dev.symbols - 1

        class _Expanding:  # the rest of the test sees images
            def __init__(self, env):
                self.env = env

            def __getattr__(self, name):
                return getattr(self.env, name)

            def step(self, a):
                o, r, d, i = self.env.step(a)
                return self.env.expand(o), r, d, i

            def step_keys(self, k):
                o, r, d, i = self.env.step_keys(k)
                return self.env.expand(o), r, d, i

        obs = dev.expand(obs)
        dev = _Expanding(dev)
    assert obs.shape == (n,) + dev.observation_space.shape and obs.is_cuda
    assert np.array_equal(obs.cpu().numpy(), ref.game.encode_states(states, *setting.encoder_args()))
    key_to_index = {ord(k): i for i, k in enumerate(envs.RogueEnv.ACTIONS)}
    stream = torch.cuda.Stream()
    for t in range(steps):
        a = np.array([key_to_index[int(k)] for k in oracle.synthetic_actions(t, seeds)], np.int64)
        states, rewards, dones, _ = ref.step(a.tolist())
        with torch.cuda.stream(stream):  # a side stream: the env must order itself against it
            obs, reward, done, _ = dev.step(torch.from_numpy(a).to("cuda", non_blocking=True))
            got_obs, got_r, got_d = obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy()
        want = ref.game.encode_states(states, *setting.encoder_args())
        assert np.array_equal(got_obs, want), "image differs at step %d" % t
        assert np.array_equal(got_r, np.asarray(rewards, np.float32)), "reward differs at step %d" % t
        assert np.array_equal(got_d.astype(bool), np.asarray(dones)), "done differs at step %d" % t
    assert not dev.errors().any()
    # ASCII keys (capitals = MoveUntil) go through the same fused call; an index outside the table only marks the env
    keys = np.frombuffer(b"HJKLhjkl.s>", np.uint8)[np.arange(n) % 11]
    states, rewards, dones, _ = ref.step("".join(chr(k) for k in keys))
    obs, reward, done, _ = dev.step_keys(keys)
    assert np.array_equal(obs.cpu().numpy(), ref.game.encode_states(states, *setting.encoder_args()))
    assert np.array_equal(reward.cpu().numpy(), np.asarray(rewards, np.float32))
    before = dev.screen.cpu().numpy().copy()
    bad = np.zeros(n, np.int64)
    bad[0] = 99
    dev.step(bad)
    assert dev.errors()[0] == 1 and np.array_equal(dev.screen.cpu().numpy()[0], before[0])
    ref.step([0] * n)
    states = ref.states
    scr = dev.screen.cpu().numpy()
    assert scr.shape == (n, 24, 80) and (scr == ord("@")).sum(axis=(1, 2)).max() == 1
    hist = dev.history().cpu().numpy()
    assert hist.shape == (n, 24, 80) and set(np.unique(hist)) <= {0, 1}
    want_hist = np.stack([s._history for s in states]).reshape(n, 24, 80)
    assert np.array_equal(hist, want_hist)
    dev.close()
    ref.close()
    with pytest.raises(RuntimeError):
        dev.step(np.zeros(n, np.int64))
