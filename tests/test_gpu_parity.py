"""Parity of the CUDA path with the CPU oracle, through the C ABI (include/rogue_b200.h) and
its Python mirror. Bit-exact: every quantity on this path is integer / byte / index work; the
f32 encoders write exactly representable values except gray = sym / symbols (one IEEE f32
divide on both sides, compared exactly as well).
"""
import json

import numpy as np
import pytest

from helpers import CONFIGS
from helpers import KEYS19, diff_dumps, diff_obs, gpu_dump, oracle_dump

pytestmark = pytest.mark.gpu


def make_pair(gpu, oracle, cfg, n, seeds, max_steps=1000):
    cfgs = json.dumps(cfg)
    pg = gpu.ParallelGameState(max_steps, [cfgs] * n)
    pg.seed([int(s) for s in seeds])
    pg.reset()
    ob = oracle.OracleBatch(cfg, n, max_steps=max_steps, seeds=[int(s) for s in seeds])
    ob.reset()
    return pg, ob


def step_gpu(b, keys):
    """rg_step_host with auto-reset; a sticky reference-panic env makes every
    batch call raise like the reference's dead worker does - tolerated here, compared separately."""
    try:
        b.step(keys, True)
    except RuntimeError as e:
        if getattr(e, "code", None) != 3:
            raise


def live_mask(pg, ob):
    """Envs in which neither side has hit a reference panic (their state is undefined afterwards)."""
    return (ob.rc != 3) & (ob.rc != 4) & (pg._batch.error != 3) & (pg._batch.error != 4)


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_reset_parity(gpu, oracle, name):
    cfg = CONFIGS[name]
    n = 384
    seeds = np.arange(0, n)  # includes seed 0 (rand_xorshift's zero-seed substitution, unpinned)
    pg, ob = make_pair(gpu, oracle, cfg, n, seeds)
    b = pg._batch
    problems = diff_obs(b, ob.obs(), -1, b.W, (ob.rc == 0))
    for i in range(0, n, 4):
        d = diff_dumps(gpu_dump(b, i), oracle_dump(ob.envs[i]), b.W)
        if d:
            problems.append("env %d (seed %d): %s" % (i, seeds[i], "; ".join(d[:4])))
    assert ((ob.rc != 0) == (b.error != 0)).all(), "panic sets differ: oracle %s gpu %s" % (
        np.nonzero(ob.rc)[0][:8], np.nonzero(b.error)[0][:8])
    assert not problems, "\n".join(problems[:6])
    pg.close()


@pytest.mark.parametrize("name,keyset,steps", [
    ("default", 11, 600), ("default", 19, 400), ("mini", 11, 600), ("mini", 19, 300), ("mini_nomon", 19, 400),
    ("default_clear", 19, 300), ("deep", 19, 400), ("wide", 19, 150), ("odd", 19, 300), ("grid4", 19, 200),
])
def test_rollout_parity(gpu, oracle, name, keyset, steps):
    """Lock-step rollouts with the SURVEY §8d action stream (11 gym actions) or a 19-key variant that
    adds the MoveUntil capitals; observations are compared every step, full internal state
    (grid, rooms, monsters, items, RNG streams, DistCache incl. maps) every 50 steps."""
    cfg = CONFIGS[name]
    n = 192
    seeds = np.arange(1, n + 1) * 7919
    max_steps = 120 if name in ("mini", "mini_nomon") else 250
    pg, ob = make_pair(gpu, oracle, cfg, n, seeds, max_steps=max_steps)
    b = pg._batch
    ids = np.arange(n)
    problems = []
    gold_before = b.status[:, 1].astype(np.int64).copy()
    for t in range(steps):
        if keyset == 11:
            keys = oracle.synthetic_actions(t, ids)
        else:
            idx = (np.frombuffer(oracle.synthetic_actions(t, ids * 31 + 7).tobytes(), np.uint8).astype(np.int64)
                   + t * 5 + ids * 3) % 19
            keys = KEYS19[idx]
        step_gpu(b, keys)
        ob.step(keys, True)
        live = live_mask(pg, ob)
        problems += diff_obs(b, ob.obs(), t, b.W, live)
        # reward = max(0, displayed gold after - before), as ParallelRogueEnv.step (parallel.py:60-63)
        want = np.maximum(0, b.status[:, 1].astype(np.int64) - gold_before)
        if not np.array_equal(b.reward[live], want[live]):
            problems.append("step %d: reward mismatch" % t)
        gold_before = b.status[:, 1].astype(np.int64).copy()
        dead_o, dead_g = (ob.rc == 3), (b.error == 3)
        if not np.array_equal(dead_o, dead_g):
            problems.append("step %d: panic sets differ: oracle %s gpu %s" % (t, np.nonzero(dead_o)[0][:8], np.nonzero(dead_g)[0][:8]))
        if t % 50 == 49 or problems:
            for i in range(0, n, 6):
                d = diff_dumps(gpu_dump(b, i), oracle_dump(ob.envs[i]), b.W)
                if d:
                    problems.append("step %d env %d: %s" % (t, i, "; ".join(d[:5])))
        if problems:
            break
    assert not problems, "\n".join(problems[:8])
    hashes = np.zeros(n, np.uint64)
    from rogue_gym_python import _cabi
    _cabi.check(b.L.rg_state_hash(b.h, hashes.ctypes.data), b.h)
    live = live_mask(pg, ob)
    assert np.array_equal(hashes[live], ob.hashes()[live])
    assert live.sum() > n // 2
    pg.close()


def test_reference_goldens_through_the_api(gpu, fixtures):
    """The reference's own live known answers, through GameState exactly as its tests drive it."""
    fx = fixtures["seed1_dungeon_clear"]
    g = gpu.GameState(1000, json.dumps(fx["config"]))
    assert g.prev().dungeon == fx["screen"]
    assert g.screen_size() == (24, 80) and g.symbols() == 17
    # test_ff_env.py: FirstFloorEnv reward 102, done on level 2, image (18, 24, 80), config round trip
    fx = fixtures["first_floor"]
    g = gpu.GameState(1000, json.dumps(fx["config"]))
    gold0 = g.prev().gold
    for k in fx["keys"]:
        g.react(ord(k))
    st = g.prev()
    assert st.gold - gold0 + (fx["stair_reward"] if st.dungeon_level > 1 else 0) == fx["expect_reward"]
    assert st.dungeon_level == 2
    assert list(st.symbol_image(fx["image_status_flag"]).shape) == fx["expect_image_shape"]
    assert json.loads(g.dump_config()) == fx["config"]
    hist = json.loads(g.dump_history())
    assert len(hist) == len(fx["keys"]) and hist[0] == {"Act": {"Move": "Up"}} and hist[-1] == {"Act": "DownStair"}
    # test_st_env.py: StairRewardEnv 104 then 100 over three generated levels, image planes, status vector
    fx = fixtures["stair_reward"]
    g = gpu.GameState(1000, json.dumps(fx["config"]))
    level, rewards = 1, []
    for keys in fx["keys"]:
        gold0 = g.prev().gold
        for k in keys:
            g.react(ord(k))
        st = g.prev()
        r = st.gold - gold0
        if st.dungeon_level > level:
            level, r = st.dungeon_level, r + fx["stair_reward"]
        rewards.append(r)
    assert rewards == fx["expect_rewards"]
    img = st.symbol_image_with_hist(fx["image_status_flag"])
    assert list(img.shape) == fx["expect_image_shape"]
    assert img[17][0][0] == fx["expect_img_17_0_0"] and img[18][0][0] == fx["expect_img_18_0_0"]
    assert st.status_vec(0x1FF) == fx["expect_full_status_vec"]


def test_reference_screens_with_monsters(gpu, fixtures):
    """The reference's known answers with monsters enabled (python/tests/data.py SEED1_DUNGEON2 / SEED1_DUNGEON3, default
    config, seed 1), through GameState.react like test_rogue_env.py::test_action drives them."""
    fx = fixtures["seed1_with_monsters"]
    for case in fx["cases"]:
        g = gpu.GameState(1000, json.dumps(fx["config"]))
        for k in case["keys"]:
            g.react(ord(k))
        assert g.prev().dungeon == case["screen"], case["keys"]


def test_reference_recordings_through_the_api(gpu, fixtures, gif_frames):
    """The reference's rendered recordings (data/gif/*.gif, decoded into tests/golden/reference_gif_frames.json): 28 + 48
    frames of two real runs of the 32x16 game, reproduced frame by frame through GameState on the GPU - dungeon rows,
    status line, messages (see tests/test_oracle_golden.py::test_reference_recordings_frame_by_frame)."""
    from helpers import check_against_recording
    from test_oracle_golden import _ddqn_keys
    for name, keys in (("ddqn_small", _ddqn_keys(fixtures, gif_frames["ddqn_small"]["n_actions"])),
                       ("ppo_cog19", gif_frames["ppo_cog19"]["keys"].encode())):
        frames = gif_frames[name]["frames"]
        g = gpu.GameState(2000, json.dumps(gif_frames["config"]))
        b = g._batch
        n = check_against_recording(g.react, lambda: g.prev().dungeon, lambda: b.status[0], lambda: b.message[0], keys, frames)
        assert n == len(frames), (name, n, len(frames))


def test_move_enemy_known_answer(gpu, fixtures):
    """core/src/dungeon/rogue/mod.rs:566-578: BFS chase with the Right/RightDown tie."""
    import ctypes as C
    fx = fixtures["move_enemy"]
    g = gpu.GameState(1000, json.dumps(fx["config"]))
    b = g._batch
    kind, nx, ny = C.c_int(), C.c_int(), C.c_int()
    from rogue_gym_python import _cabi
    _cabi.check(b.L.rg_test_move_enemy(b.h, 0, fx["from"][0], fx["from"][1], fx["to"][0], fx["to"][1], C.byref(kind),
                                       C.byref(nx), C.byref(ny)), b.h)
    assert (kind.value, nx.value, ny.value) == (1, fx["expect"][0], fx["expect"][1])


@pytest.mark.parametrize("name", ["default", "mini", "odd"])
def test_encoders_match_oracle(gpu, oracle, name):
    """rg_encode over the whole resident batch and rg_encode_states on detached PlayerStates."""
    import ctypes as C
    import torch
    cfg = CONFIGS[name]
    n = 48
    seeds = np.arange(1, n + 1)
    pg, ob = make_pair(gpu, oracle, cfg, n, seeds)
    b = pg._batch
    ids = np.arange(n)
    for t in range(60):
        keys = oracle.synthetic_actions(t, ids)
        step_gpu(b, keys)
        ob.step(keys, True)
    from rogue_gym_python import _cabi
    states = pg.states()
    for mode, flag, hist in [(0, 0, 0), (0, 0x1FF, 1), (1, 0, 0), (1, 0x1FF, 0), (1, 0x83, 1), (0, 0x10, 0)]:
        ch = b.L.rg_encode_channels(b.h, mode, flag, hist)
        out = torch.empty((n, ch, b.H, b.W), dtype=torch.float32, device="cuda")
        got_ch = C.c_int()
        rc = b.L.rg_encode(b.h, mode, flag, hist, out.data_ptr(), C.byref(got_ch))
        _cabi.check(rc, b.h)
        rc = b.L.rg_sync(b.h)
        assert rc in (0, 3, 4)
        got = out.cpu().numpy()
        assert got_ch.value == ch
        for i in range(n):
            if ob.rc[i] != 0 or b.error[i] == 3:
                continue
            try:
                want = ob.envs[i].encode(mode, flag, bool(hist))
            except oracle.OracleError:
                continue  # InvalidTileError ('Z' with the default table): flagged on the device
            assert got[i].shape == want.shape
            assert np.array_equal(got[i], want), "mode %d flag %x hist %d env %d" % (mode, flag, hist, i)
            if i < 4:
                fn = {(0, 0): "gray_image", (0, 1): "gray_image_with_hist", (1, 0): "symbol_image",
                      (1, 1): "symbol_image_with_hist"}[(mode, hist)]
                assert np.array_equal(getattr(states[i], fn)(flag), want)
    pg.close()


def test_invalid_tile_error(gpu):
    """symbol.rs:60-64: a visible tile whose symbol id >= symbols-1 raises (reachable with a
    monster table whose largest tile is on screen)."""
    cfg = {"seed": 3, "hide_dungeon": False,
           "enemies": {"enemies": [4], "appear_rate_gold": 100, "appear_rate_nogold": 100}}
    g = gpu.GameState(100, json.dumps(cfg))
    st = g.prev()
    if any("E" in row for row in st.dungeon):
        with pytest.raises(RuntimeError, match="Invalid tile"):
            st.symbol_image()
        assert st.gray_image().shape == (1, 24, 80)


def test_single_env_semantics(gpu, oracle):
    """python/tests/test_rogue_env.py + state_impls.rs: NoOp, max_steps, errors, seeding."""
    cfg = json.dumps({"seed": 1})
    g = gpu.GameState(5, cfg)
    before = g.prev()
    assert not before.is_terminal
    for _ in range(5):
        assert not g.prev().is_terminal
        g.react(ord("."))
    after = g.prev()
    assert after.is_terminal and after.dungeon == before.dungeon and after != before
    g.react(ord("h"))       # steps == max_steps: still processed (strict `>` early-out)
    moved = g.prev()
    g.react(ord("h"))       # now a no-op
    assert g.prev() == moved
    with pytest.raises(RuntimeError, match="Invliad input key"):
        gpu.GameState(5, cfg).react(ord("x"))
    # set_seed takes effect at the next reset and sticks
    g = gpu.GameState(100, cfg)
    first = g.prev().dungeon
    g.set_seed(7)
    assert g.prev().dungeon == first
    g.reset()
    seven = g.prev().dungeon
    assert seven == oracle.OracleEnv({"seed": 7}).dungeon() and seven != first
    g.reset()
    assert g.prev().dungeon == seven
    assert "Level:  1 Gold:     0 Hp: 12(12) Str: 16(16) Arm:  0 Exp:  1/ 0" in repr(g.prev())
    assert g.prev().status == {"dungeon_level": 1, "gold": 0, "hp_current": 12, "hp_max": 12, "str_current": 16,
                               "str_max": 16, "defense": 0, "player_level": 1, "exp": 0, "hunger": 0}


def test_death_then_ignored_input(gpu, oracle):
    """After the grave modal any action is IgnoredInput (core/src/lib.rs:314)."""
    cfg = {"seed": 11, "player": {"init_hp": 1}}
    o = oracle.OracleEnv(cfg, max_steps=5000)
    g = gpu.GameState(5000, json.dumps(cfg))
    ids = np.array([0])
    for t in range(4000):
        k = int(oracle.synthetic_actions(t, ids)[0])
        try:
            o.react(k)
            oerr = 0
        except oracle.OracleError as e:
            oerr = e.code
        try:
            g.react(k)
            gerr = 0
        except RuntimeError as e:
            gerr = e.code
        assert oerr == gerr, t
        if oerr == 2:
            break
        assert g.prev().is_terminal == o.obs()["is_terminal"]
    assert oerr == 2, "the walk never died; pick another seed"


def test_parallel_semantics(gpu, oracle):
    """python/tests/test_parallel.py + thread_impls.rs: same config => same states, auto-reset
    returns the fresh state flagged terminal, seed() applies from the next reset."""
    cfg = {"seed": 5, "width": 32, "height": 16, "dungeon": {"style": "rogue", "room_num_x": 2, "room_num_y": 2}}
    n = 8
    pg = gpu.ParallelGameState(4, [json.dumps(cfg)] * n)
    states = pg.states()
    assert all(s == states[0] for s in states) and pg.screen_size() == (16, 32) and pg.symbols() == 43
    out = pg.step([ord(c) for c in "hjklyubn"])
    assert not all(s == out[0] for s in out)
    for _ in range(2):
        out = pg.step([ord("h")] * n)
        assert not any(s.is_terminal for s in out)
    out = pg.step([ord("h")] * n)
    assert all(s.is_terminal for s in out)
    assert all(s.dungeon == states[0].dungeon for s in out)
    # only the copy that step() returned is flagged: the workers' own states are the fresh games
    # (ThreadConductor::step flags `*res`, Instruction::State returns game_state.state(): thread_impls.rs:69-79,131)
    again = pg.states()
    assert not any(s.is_terminal for s in again) and all(s.dungeon == states[0].dungeon for s in again)
    # fewer seeds than workers: zip() seeds the first ones, the others keep theirs (thread_impls.rs:45-50)
    pg.seed([100, 101])
    part = pg.reset()
    assert part[0].dungeon == oracle.OracleEnv(dict(cfg, seed=100)).dungeon()
    assert part[1].dungeon == oracle.OracleEnv(dict(cfg, seed=101)).dungeon()
    assert all(s.dungeon == states[0].dungeon for s in part[2:])
    pg.seed(list(range(100, 100 + n)))
    fresh = pg.reset()
    for i, s in enumerate(fresh):
        assert s.dungeon == oracle.OracleEnv(dict(cfg, seed=100 + i)).dungeon() and not s.is_terminal
    # per-env configs that differ in the seed only
    pg2 = gpu.ParallelGameState(10, [json.dumps(dict(cfg, seed=s)) for s in (1, 2, 3)])
    for s, st in zip((1, 2, 3), pg2.states()):
        assert st.dungeon == oracle.OracleEnv(dict(cfg, seed=s)).dungeon()
    with pytest.raises(RuntimeError, match="must agree on width"):
        gpu.ParallelGameState(10, [json.dumps(cfg), json.dumps(dict(cfg, width=40))])
    pg.close()
    with pytest.raises(RuntimeError, match="closed"):
        pg.states()


def test_unseeded_configs_draw_fresh_seeds(gpu):
    """seed: null => every env and every episode is a different game (core/src/lib.rs:157-165)."""
    pg = gpu.ParallelGameState(3, ["{}"] * 16)
    first = [tuple(s.dungeon) for s in pg.states()]
    assert len(set(first)) > 8
    for _ in range(3):
        out = pg.step([ord(".")] * 16)
    second = [tuple(s.dungeon) for s in out]
    assert all(s.is_terminal for s in out) and sum(a != b for a, b in zip(first, second)) > 8
    r = gpu.ParallelGameState(3, [json.dumps({"seed_range": [5, 6]})] * 4)
    assert len({tuple(s.dungeon) for s in r.states()}) == 1


def test_full_size_batch_properties(gpu, oracle):
    """BASELINE config 3 size (65 536 envs, default 80x24): the oracle checks a strided sample
    bit-exactly; the whole batch is checked through size-independent properties: determinism
    of a re-run, agreement of envs that share a seed, and screen/status invariants."""
    import ctypes as C
    from rogue_gym_python import _cabi
    n, steps = 65536, 60
    L = _cabi.lib()
    cfgs = (C.c_char_p * 1)(b"{}")
    seeds = (np.arange(n, dtype=np.uint64) % np.uint64(60000)) + np.uint64(1)  # envs i and i+60000 share a seed
    ids = np.arange(n)
    digests = []
    sample = np.arange(0, n, 257)
    ob = oracle.OracleBatch({}, len(sample), seeds=[int(seeds[i]) for i in sample])
    ob.reset()
    for rep in range(2):
        h = C.c_void_p()
        _cabi.check(L.rg_create(cfgs, 1, n, 1000, 0, C.byref(h)))
        _cabi.check(L.rg_seed(h, seeds.ctypes.data, None), h)
        _cabi.check(L.rg_reset(h), h)
        acts = np.zeros(n, np.uint8)
        screen = np.zeros((n, 1920), np.uint8)
        status = np.zeros((n, 10), np.uint32)
        err = np.zeros(n, np.uint8)
        obs = _cabi.HostObs(screen.ctypes.data, None, status.ctypes.data, None, None, None, err.ctypes.data)
        for t in range(steps):
            # same key for envs that share a seed
            acts[:] = oracle.synthetic_actions(t, seeds)
            rc = L.rg_step_host(h, acts.ctypes.data, 1, C.byref(obs))
            assert rc in (0, 3), L.rg_last_error(h)
            if rep == 0:
                ob.step(acts[sample], True)
        hashes = np.zeros(n, np.uint64)
        _cabi.check(L.rg_state_hash(h, hashes.ctypes.data), h)
        digests.append(hashes)
        if rep == 0:
            live = (ob.rc == 0) & (err[sample] == 0)
            assert np.array_equal(hashes[sample][live], ob.hashes()[live])
            assert np.array_equal((ob.rc == 3), (err[sample] == 3))
            assert live.sum() > 0.9 * len(sample)
            ok = err == 0
            assert np.array_equal(hashes[:5536][ok[:5536]], hashes[60000:][ok[:5536]])
            at = (screen == ord("@")).sum(axis=1)[ok]
            # one player per screen; none only when the player stands on a hidden maze cell (never `approached`)
            assert (at <= 1).all() and (at == 1).mean() > 0.995
            assert (screen[:, :80] == 32).all() and (screen[:, -80:] == 32).all()  # rows 0 and H-1 are never drawn
            assert (status[ok][:, 2] <= status[ok][:, 3]).all() and (status[ok][:, 4] == 16).all()
        L.rg_destroy(h)
    assert np.array_equal(digests[0], digests[1])


# ---------------------------------------------------------------- host mirror (delta write-back)
@pytest.mark.gpu
@pytest.mark.parametrize("cfg_name", ["default", "mini", "wide", "odd", "default-direct", "mini-direct"])
def test_host_mirror_equals_full_copy(gpu, cfg_name, monkeypatch):
    """rg_step_mirror must leave the host mirror byte-identical to what rg_fetch copies, every step,
    across auto-resets, stair descents and explicit resets, while moving far fewer bytes.
    default / mini / wide (160x48: 120 lines and 60 visited-map pieces per env) go through k_mirror_lines (whole
    64-byte lines, on a few SMs), "odd" (a screen that is not a whole
    number of 64-byte lines) and "-direct" (RG_MIRROR_MODE=direct) through k_mirror."""
    import json

    from helpers import CONFIGS
    from helpers import KEYS19
    n, steps = 192, 150
    monkeypatch.setenv("RG_MIRROR_MODE", (cfg_name.split("-") + ["lines"])[1])
    cfg = CONFIGS[cfg_name.split("-")[0]]
    pg = gpu.ParallelGameState(25, [json.dumps(cfg)] * n)
    pg.seed(list(range(3, n + 3)))
    pg.reset()
    b = pg._batch
    m = b.mirror()
    assert m["screen"].shape == (n, b.H, b.W) and not m["screen"].flags.owndata
    rng = np.random.RandomState(11)
    full_bytes = n * (b.C + 40 + 4 + 1 + 4 + 1)
    sent = []
    for t in range(steps):
        keys = KEYS19[rng.randint(0, len(KEYS19), size=n)]
        if t % 40 == 39:  # a step that is not mirrored: the next mirrored step has to bring its changes along
            try:
                b.step(KEYS19[rng.randint(0, len(KEYS19), size=n)], True)
            except RuntimeError as e:
                assert getattr(e, "code", None) in (2, 3), e  # per-env errors (input after death, reference panic)
        sent.append(b.step_mirror(keys, True))
        b.fetch()  # the full D2H copy of the same block
        assert np.array_equal(m["screen"].reshape(n, -1), b.screen), "screen differs at step %d" % t
        assert np.array_equal(b.mirror_history().reshape(n, -1), b.history), "history differs at step %d" % t
        assert np.array_equal(m["status"], b.status) and np.array_equal(m["reward"], b.reward)
        assert np.array_equal(m["done"], b.done) and np.array_equal(m["message"], b.message)
        assert np.array_equal(m["error"], b.error)
        if t == 70:
            pg.seed(list(range(1000, n + 1000)))
            b.reset()
            assert b.mirror_sync() > 0
            assert np.array_equal(m["screen"].reshape(n, -1), b.screen)
    assert b.mirror_sync() == 0  # nothing changed since the last step
    assert np.median(sent) < 0.25 * full_bytes, (np.median(sent), full_bytes)
    arr = pg.step_arrays(KEYS19[rng.randint(0, 11, size=n)])
    b.fetch()
    assert np.array_equal(arr["screen"].reshape(n, -1), b.screen) and np.array_equal(arr["done"], b.done)
    pg.close()


# ---------------------------------------------------------------- heterogeneous batches
@pytest.mark.gpu
def test_per_env_configs(gpu, oracle):
    """Every worker of the reference's ParallelGameState is built from its own JSON (python/src/lib.rs:
    270-280). One batch here: four different configs with the same geometry, interleaved; every env must
    follow the oracle built from ITS config, step by step."""
    import json

    from helpers import KEYS19, diff_obs
    variants = [
        {},
        {"enemies": {"enemies": [0, 3, 7, 12, 25]}, "hide_dungeon": False},
        {"dungeon": {"style": "rogue", "dark_level": 1, "maze_rate_inv": 3, "locked_door_rate_inv": 2,
                     "hidden_passage_rate_inv": 5},
         "item": {"gold": {"rate_inv": 1, "base": 200, "per_level": 50, "minimum": 10}}},
        {"player": {"init_hp": 40, "hunger_time": 60}, "enemies": {"enemies": list(range(26)), "appear_rate_gold": 100,
                                                                  "appear_rate_nogold": 100}},
    ]
    n, steps = 128, 160
    cfgs = [variants[i % 4] for i in range(n)]
    seeds = list(range(11, n + 11))
    pg = gpu.ParallelGameState(60, [json.dumps(c) for c in cfgs])
    pg.seed(seeds)
    pg.reset()
    b = pg._batch
    obs = [oracle.OracleBatch(variants[k], n // 4, max_steps=60, seeds=seeds[k::4]) for k in range(4)]
    for ob in obs:
        ob.reset()
    rng = np.random.RandomState(5)

    def gather():
        parts = [ob.obs() for ob in obs]
        out = {}
        for key in parts[0]:
            full = np.zeros((n,) + parts[0][key].shape[1:], parts[0][key].dtype)
            for k in range(4):
                full[k::4] = parts[k][key]
            out[key] = full
        rc = np.zeros(n, np.int32)
        for k in range(4):
            rc[k::4] = obs[k].rc
        return out, rc

    for t in range(steps):
        keys = KEYS19[rng.randint(0, len(KEYS19), size=n)]
        try:
            b.step(keys, True)
        except RuntimeError as e:
            assert getattr(e, "code", None) == 3  # a reference-panic state somewhere: sticky, compared below
        for k in range(4):
            obs[k].step(keys[k::4], True)
        want, rc = gather()
        live = (rc == 0) & (b.error == 0)
        assert np.array_equal(rc == 3, b.error == 3), "panic sets differ at step %d" % t
        bad = diff_obs(b, want, t, b.W, live)
        assert not bad, "\n".join(bad)
    assert live.sum() > n // 2
    # variants are really different games
    assert len({bytes(b.screen[i]) for i in range(4)}) > 1
    pg.close()
    with pytest.raises(RuntimeError, match="must agree on width"):
        gpu.ParallelGameState(60, [json.dumps({}), json.dumps({"width": 64})])


# ---------------------------------------------------------------- sharding invariance
@pytest.mark.gpu
def test_shards_reproduce_the_single_batch(gpu):
    """An env's trajectory depends only on its id (seed, action stream), never on which shard or batch
    size it is stepped in: two shards of 160 envs give the state hashes of one batch of 320."""
    from rogue_gym_python.rollout import Shard, synthetic_actions
    n, steps = 320, 120
    whole = Shard("{}", 0, n, max_steps=50)
    parts = [Shard("{}", 0, n // 2, max_steps=50), Shard("{}", n // 2, n, max_steps=50)]
    import torch
    for t in range(steps):
        for sh in [whole] + parts:
            keys = torch.from_numpy(synthetic_actions(t, sh.env_ids)).cuda()
            sh.step_device(keys.data_ptr())
            sh.sync()
    h = whole.hashes()
    assert np.array_equal(h[: n // 2], parts[0].hashes()) and np.array_equal(h[n // 2:], parts[1].hashes())
    assert np.array_equal(whole.errors()[n // 2:], parts[1].errors())
    for sh in [whole] + parts:
        sh.close()


# ---------------------------------------------------------------- panic policy
@pytest.mark.gpu
def test_panic_policy_terminal(gpu, oracle):
    """A state in which the reference panics (monster at x = 0 probing x = -1, rogue/mod.rs:361) kills the reference's
    worker and with it the conductor (thread_impls.rs:111-135). Default here ("sticky", covered by every other test):
    the env is frozen. With the "terminal" policy the step that hits the state reports done = 1 and error 3 once and
    the env continues with a fresh episode; every other env and every other step is untouched."""
    n, steps, max_steps = 2048, 150, 400
    seeds = np.arange(1, n + 1)  # sequential seeds: ~1.5 % of them reach the panic state within a few steps
    pg, ob = make_pair(gpu, oracle, {}, n, seeds, max_steps=max_steps)
    b = pg._batch
    b.set_panic_policy("terminal")
    ids = np.arange(n)
    revived_total = 0
    born_dead = ob.rc == 3  # the seeded game itself starts in the panic state: it ends at its first step
    for t in range(steps):
        keys = oracle.synthetic_actions(t, ids)
        try:
            b.step(keys, True)
        except RuntimeError as e:
            assert getattr(e, "code", None) == 3
        ob.step(keys, True)
        hit = (ob.rc == 3)
        assert np.array_equal(b.error == 3, hit), "step %d: panic events differ: gpu %s oracle %s" % (
            t, np.nonzero(b.error == 3)[0][:8], np.nonzero(hit)[0][:8])
        assert (b.done[hit] == 1).all()
        for i in np.nonzero(hit)[0]:  # what the policy does, on the oracle: a fresh episode of the same seed
            try:
                ob.envs[i].reset()
            except oracle.OracleError:
                born_dead[i] = True  # the seed's game panics while it is generated: it ends again at every step
        ob.rc[hit] = 0
        o = ob.obs()
        o["done"][hit] = 1
        bad = diff_obs(b, o, t, b.W, ~born_dead)
        assert not bad, "\n".join(bad[:4])
        revived_total += int((hit & ~born_dead).sum())
    assert revived_total >= 10 and born_dead.sum() < n // 100
    hashes = np.zeros(n, np.uint64)
    from rogue_gym_python import _cabi
    _cabi.check(b.L.rg_state_hash(b.h, hashes.ctypes.data), b.h)
    assert np.array_equal(hashes[~born_dead], ob.hashes()[~born_dead])
    pg.close()
