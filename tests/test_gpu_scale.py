"""Parity at BASELINE.json's full sizes and through the rare paths that only occur at scale:

* 65 536 default envs stepped past step 1 000 (every env ends its first episode on the same step: the
  mass swap-in of prefetched games and the refill burst behind it), an oracle sample of 512 envs
  compared by state hash every 100 steps and by screen / status at the end;
* a ring-pressure run (max_steps 2-3: an env ends episodes faster than the background generator refills
  its two-slot ring), which forces the synchronous-reset and stale-slot paths, checked bit-exactly on
  every env;
* the reference's own generator property tests restated on > 1 M GPU-generated floors of three sizes
  (passages.rs:343-379 `connectivity`, rooms.rs:308-340 `pos_check`, floor.rs:466-488 `secret_door`);
* the shipped data/config-default.json and BASELINE.json configs[0] (config-mini, seeds 0 and 4,
  200 scripted keys incl. the MoveUntil capitals) rolled out step by step.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from helpers import ROOT
from helpers import KEYS19, diff_dumps, diff_obs, gpu_dump, oracle_dump

pytestmark = pytest.mark.gpu


def _raw_batch(cabi, cfg, n, max_steps, seeds):
    L = cabi.lib()
    cfgs = (C.c_char_p * 1)(json.dumps(cfg).encode())
    h = C.c_void_p()
    cabi.check(L.rg_create(cfgs, 1, n, max_steps, 0, C.byref(h)))
    seeds = np.ascontiguousarray(seeds, np.uint64)
    cabi.check(L.rg_seed(h, seeds.ctypes.data, None), h)
    cabi.check(L.rg_reset(h), h)
    return L, h


def _stats(L, h, cabi):
    out = np.zeros(8, np.uint64)
    cabi.check(L.rg_stats(h, out.ctypes.data), h)
    names = ("swap_in", "sync_reset", "full_step", "prefetch_built", "prefetch_stale", "monster_env_steps", "bfs_levels",
             "fast_steps")
    return {k: int(v) for k, v in zip(names, out)}


def test_full_size_past_the_mass_reset(gpu, cabi, oracle):
    """BASELINE configs[2]: 65 536 default envs, seeds 1+i, the bench's action stream, 1 150 steps."""
    import torch
    n, steps, max_steps = 65536, 1150, 1000
    seeds = np.arange(n, dtype=np.uint64) + np.uint64(1)
    L, h = _raw_batch(cabi, {}, n, max_steps, seeds)
    sample = np.arange(0, n, 128)  # 512 envs
    ob = oracle.OracleBatch({}, len(sample), max_steps=max_steps, seeds=[int(seeds[i]) for i in sample])
    ob.reset()
    ids = np.arange(n, dtype=np.uint64)
    hashes = np.zeros(n, np.uint64)
    err = np.zeros(n, np.uint8)
    eobs = cabi.HostObs(None, None, None, None, None, None, err.ctypes.data)
    # rg_step is asynchronous on the batch's own stream: every step gets its own resident action row
    all_keys = np.stack([oracle.synthetic_actions(t, ids) for t in range(steps)])
    dev = torch.from_numpy(all_keys).cuda()
    torch.cuda.synchronize()
    checked = 0
    for t in range(steps):
        keys = all_keys[t]
        rc = L.rg_step(h, dev[t].data_ptr(), 1)
        assert rc == 0, L.rg_last_error(h)
        ob.step(keys[sample], True)
        if t % 100 == 99 or t in (998, 999, 1000, 1001, 1002, steps - 1):
            cabi.check(L.rg_state_hash(h, hashes.ctypes.data), h)
            cabi.check(L.rg_fetch(h, C.byref(eobs)), h)
            live = (ob.rc == 0) & (err[sample] == 0)
            assert np.array_equal(ob.rc == 3, err[sample] == 3), "panic sets differ at step %d" % t
            bad = np.nonzero(hashes[sample][live] != ob.hashes()[live])[0]
            assert len(bad) == 0, "step %d: %d sampled envs differ from the oracle, first env %d" % (
                t, len(bad), int(sample[live][bad[0]]))
            checked += 1
    assert checked >= 15 and live.sum() > 0.9 * len(sample)
    # the end state, observation by observation
    screen = np.zeros((n, 1920), np.uint8)
    status = np.zeros((n, 10), np.uint32)
    done = np.zeros(n, np.uint8)
    obs = cabi.HostObs(screen.ctypes.data, None, status.ctypes.data, None, done.ctypes.data, None, None)
    cabi.check(L.rg_fetch(h, C.byref(obs)), h)
    o = ob.obs()
    assert np.array_equal(screen[sample][live], o["screen"][live])
    assert np.array_equal(status[sample][live], o["status"][live])
    assert np.array_equal(done[sample][live], o["done"][live])
    st = _stats(L, h, cabi)
    # every live env ended its first episode at step 1000 and took a prefetched game (or a synchronous one)
    assert st["swap_in"] + st["sync_reset"] >= int((err == 0).sum()), st
    assert st["swap_in"] > 0.99 * (st["swap_in"] + st["sync_reset"]), st
    assert st["prefetch_built"] >= st["swap_in"], st
    L.rg_destroy(h)


@pytest.mark.parametrize("max_steps,n", [(2, 4096), (3, 65536)])
def test_ring_pressure_hits_the_miss_paths(gpu, cabi, oracle, max_steps, n, monkeypatch):
    """Episodes of 2-3 steps with a background pass only every 8th step drain every env's two-slot ring faster
    than it is refilled: finish_env's miss path (synchronous reset in k_step_gen), the cancel marks and the
    stale-slot check all run. Whatever path an episode's game comes from, it must be the game the oracle builds."""
    import torch
    monkeypatch.setenv("RG_PREFETCH_EVERY", "8")  # read by rg_create
    steps = 120 if n <= 4096 else 40
    seeds = (np.arange(n, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(17)) % np.uint64(1 << 40)
    L, h = _raw_batch(cabi, {}, n, max_steps, seeds)
    sample = np.arange(n) if n <= 4096 else np.arange(0, n, 64)
    ob = oracle.OracleBatch({}, len(sample), max_steps=max_steps, seeds=[int(seeds[i]) for i in sample])
    ob.reset()
    ids = np.arange(n, dtype=np.uint64)
    hashes = np.zeros(n, np.uint64)
    err = np.zeros(n, np.uint8)
    eobs = cabi.HostObs(None, None, None, None, None, None, err.ctypes.data)
    all_keys = np.stack([oracle.synthetic_actions(t, ids) for t in range(steps)])
    dev = torch.from_numpy(all_keys).cuda()  # one resident row per step (rg_step is asynchronous)
    torch.cuda.synchronize()
    for t in range(steps):
        keys = all_keys[t]
        assert L.rg_step(h, dev[t].data_ptr(), 1) == 0, L.rg_last_error(h)
        ob.step(keys[sample], True)
        if t % 8 == 7 or t == steps - 1:
            cabi.check(L.rg_state_hash(h, hashes.ctypes.data), h)
            cabi.check(L.rg_fetch(h, C.byref(eobs)), h)
            live = (ob.rc == 0) & (err[sample] == 0)
            assert np.array_equal(ob.rc == 3, err[sample] == 3), "panic sets differ at step %d" % t
            assert np.array_equal(hashes[sample][live], ob.hashes()[live]), "step %d" % t
    st = _stats(L, h, cabi)
    assert st["sync_reset"] > 0, st          # the ring ran dry: games built inside the step
    assert st["swap_in"] > 0, st             # and prefetched games were used as well
    assert st["swap_in"] + st["sync_reset"] >= (steps // max_steps - 1) * int((err == 0).sum()), st
    L.rg_destroy(h)


# ------------------------------------------------------------------------------------------------
# The reference's generator property tests on whole batches of GPU floors. The planes stay on the
# device (rg_export_floors) and are checked there with torch; nothing here is on the product path.
S_PASSAGE, S_FLOOR, S_WALLX, S_WALLY, S_STAIR, S_DOOR, S_TRAP, S_NONE = range(8)
A_HIDDEN, A_LOCKED, A_DOOR = 2, 16, 64
SIZES = {
    "32x16": {"width": 32, "height": 16, "dungeon": {"style": "rogue", "room_num_x": 2, "room_num_y": 2}},
    "80x24": {},
    "160x48": {"width": 160, "height": 48},
}


def _export(L, h, cabi, n, W, H, nrooms_max=16):
    import torch
    surf = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
    attr = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
    rooms = torch.empty((n, nrooms_max, 8), dtype=torch.int16, device="cuda")
    cabi.check(L.rg_export_floors(h, surf.data_ptr(), attr.data_ptr(), rooms.data_ptr()), h)
    rc = L.rg_sync(h)
    assert rc in (0, 3, 4), L.rg_last_error(h)
    return surf, attr, rooms


def _flood_all_reached(surf, attr):
    """passages.rs:343-379: from the first Floor cell, a 4-direction flood over can_walk cells reaches every
    can_walk cell. The reference's buffer holds the dug surfaces; here a hidden passage / locked door keeps its
    old surface until found, so "walkable" = can_walk(surface) or HIDDEN or LOCKED. Returns (has_start, ok) per env."""
    import torch
    walk = ((surf != S_WALLX) & (surf != S_WALLY) & (surf != S_NONE)) | ((attr & (A_HIDDEN | A_LOCKED)) != 0)
    n, H, W = surf.shape
    floor = (surf == S_FLOOR).reshape(n, -1)
    has = floor.any(dim=1)
    first = torch.argmax(floor.to(torch.uint8), dim=1)
    vis = torch.zeros((n, H * W), dtype=torch.bool, device=surf.device)
    vis[torch.arange(n, device=surf.device)[has], first[has]] = True
    vis = vis.reshape(n, H, W)
    for it in range(4096):
        grow = vis.clone()
        grow[:, 1:, :] |= vis[:, :-1, :]
        grow[:, :-1, :] |= vis[:, 1:, :]
        grow[:, :, 1:] |= vis[:, :, :-1]
        grow[:, :, :-1] |= vis[:, :, 1:]
        grow &= walk
        if it % 16 == 15 and bool((grow == vis).all()):
            break
        vis = grow
    ok = (vis == walk).reshape(n, -1).all(dim=1)
    return has, ok


@pytest.mark.parametrize("size", sorted(SIZES))
def test_generator_properties_on_gpu_floors(gpu, cabi, size):
    """>= 1 M floors over the three sizes (BASELINE configs[4]'s sweep): 327 680 / 524 288 / 196 608."""
    import torch
    n = 65536
    reps = {"32x16": 5, "80x24": 8, "160x48": 3}[size]
    # dark rooms, mazes, hidden passages and locked doors already on level 1 (the reference test generates level 10)
    cfg = json.loads(json.dumps(SIZES[size]))
    dg = cfg.setdefault("dungeon", {"style": "rogue"})
    dg.update({"dark_level": 2, "maze_rate_inv": 3, "hidden_passage_rate_inv": 6, "locked_door_rate_inv": 3})
    W, H = cfg.get("width", 80), cfg.get("height", 24)
    nx, ny = dg.get("room_num_x", 3), dg.get("room_num_y", 3)
    L, h = _raw_batch(cabi, cfg, n, 10, np.arange(n, dtype=np.uint64) + np.uint64(1))
    err = np.zeros(n, np.uint8)
    eobs = cabi.HostObs(None, None, None, None, None, None, err.ctypes.data)
    floors = mazes = 0
    for rep in range(reps):
        if rep:
            seeds = np.arange(n, dtype=np.uint64) + np.uint64(1 + rep * n)
            cabi.check(L.rg_seed(h, seeds.ctypes.data, None), h)
            cabi.check(L.rg_reset(h), h)
        surf, attr, rooms = _export(L, h, cabi, n, W, H)
        cabi.check(L.rg_fetch(h, C.byref(eobs)), h)
        good = torch.from_numpy(err == 0).cuda()
        # connectivity
        has, ok = _flood_all_reached(surf, attr)
        bad = torch.nonzero(good & has & ~ok).flatten()
        assert bad.numel() == 0, "%s rep %d: %d floors not connected, first env %d" % (size, rep, bad.numel(), int(bad[0]))
        # pos_check: every ranged room has area >= 9 and keeps to its side of its right / lower neighbour
        r = rooms[:, : nx * ny].to(torch.int32)
        kind, x0, y0, x1, y1 = r[..., 0], r[..., 2], r[..., 3], r[..., 4], r[..., 5]
        ranged = (kind == 0) | (kind == 1)
        area_ok = (~ranged) | ((x1 - x0) * (y1 - y0) >= 9)
        assert bool(area_ok[good].all())
        for a in range(nx * ny):
            ax, ay = a % nx, a // nx
            if ax + 1 < nx:
                b = a + 1
                both = ranged[:, a] & ranged[:, b] & good
                assert bool((x0[:, b] >= x1[:, a])[both].all()), "rooms %d/%d overlap in x" % (a, b)
            if ay + 1 < ny:
                b = a + nx
                both = ranged[:, a] & ranged[:, b] & good
                assert bool((y0[:, b] >= y1[:, a])[both].all()), "rooms %d/%d overlap in y" % (a, b)
        # a floor has exactly one stair, and the player stands on a walkable cell inside the field
        assert bool(((surf == S_STAIR).reshape(n, -1).sum(dim=1) == 1)[good].all())
        px, py = rooms[:, 0, 6].long(), rooms[:, 0, 7].long()
        ps = surf[torch.arange(n, device="cuda"), py.clamp(0, H - 1), px.clamp(0, W - 1)]
        assert bool((((ps == S_FLOOR) | (ps == S_PASSAGE) | (ps == S_STAIR)) & (py >= 1) & (py < H - 1))[good].all())
        floors += int(good.sum())
        mazes += int(((kind == 1).any(dim=1) & good).sum())
    assert floors > 0.95 * n * reps and mazes > 0.2 * floors, (floors, mazes)
    L.rg_destroy(h)


def test_secret_doors_grow_with_depth(gpu, cabi):
    """floor.rs:466-488 `secret_door`: the number of door-set cells that are not (yet) doors - locked doors -
    does not fall as the level rises. gen_attr locks a door when range(0, dark_level) < level and a
    1/locked_door_rate_inv roll hit (floor.rs:420-451), so on level-1 floors dark_level = 10, 5, 3, 2, 1
    gives the lock probability of levels 1, 2, 3.3, 5 and >= 10 of the default config."""
    import torch
    n = 65536
    before = -1.0
    for dark in (10, 5, 3, 2, 1):
        cfg = {"dungeon": {"style": "rogue", "dark_level": dark}}
        L, h = _raw_batch(cabi, cfg, n, 10, np.arange(n, dtype=np.uint64) + np.uint64(7))
        surf, attr, _ = _export(L, h, cabi, n, 80, 24)
        hidden = ((attr & A_DOOR) != 0) & (surf != S_DOOR)
        assert bool((((attr & A_LOCKED) != 0) == hidden).all())  # exactly the locked ones
        per100 = float(hidden.sum()) / n * 100
        assert before <= per100 + 10, (dark, before, per100)  # the reference's slack: 10 per 100 floors
        assert per100 > before, (dark, before, per100)        # at this sample size it is strictly monotone
        before = per100
        L.rg_destroy(h)


# ------------------------------------------------------------------------------------------------
def _lockstep(gpu, oracle, cfg, n, seeds, max_steps, keys_at, steps, dump_every=50):
    cfgs = json.dumps(cfg)
    pg = gpu.ParallelGameState(max_steps, [cfgs] * n)
    pg.seed([int(s) for s in seeds])
    pg.reset()
    ob = oracle.OracleBatch(cfg, n, max_steps=max_steps, seeds=[int(s) for s in seeds])
    ob.reset()
    b = pg._batch
    problems = diff_obs(b, ob.obs(), -1, b.W, ob.rc == 0)
    for t in range(steps):
        keys = keys_at(t)
        try:
            b.step(keys, True)
        except RuntimeError as e:
            if getattr(e, "code", None) != 3:
                raise
        ob.step(keys, True)
        live = (ob.rc != 3) & (ob.rc != 4) & (b.error != 3) & (b.error != 4)
        problems += diff_obs(b, ob.obs(), t, b.W, live)
        if not np.array_equal(ob.rc == 3, b.error == 3):
            problems.append("step %d: panic sets differ" % t)
        if t % dump_every == dump_every - 1 or problems:
            for i in range(0, n, max(1, n // 32)):
                d = diff_dumps(gpu_dump(b, i), oracle_dump(ob.envs[i]), b.W)
                if d:
                    problems.append("step %d env %d: %s" % (t, i, "; ".join(d[:5])))
        if problems:
            break
    assert not problems, "\n".join(problems[:8])
    assert live.sum() > n // 2
    pg.close()


def test_shipped_default_config_rollout(gpu, oracle):
    """data/config-default.json as shipped (tests/golden/config_default.json: every field spelled out, `exps`
    ending in 0, `seed: null`), not the `{}` shorthand the other tests use."""
    with open(os.path.join(ROOT, "tests", "golden", "config_default.json")) as f:
        cfg = json.load(f)
    n = 192
    ids = np.arange(n)
    _lockstep(gpu, oracle, cfg, n, np.arange(1, n + 1) * 104729, 300,
              lambda t: oracle.synthetic_actions(t, ids), 450)


SCRIPT = ("hhhhjjjjllllkkkk" "yubn" "s.>" "HJKLYUBN" "llllllll" "jjjj" ">" "ssss" "hjklhjkl" "LKJH" "nbuy" ".")


@pytest.mark.parametrize("seed", [0, 4])
def test_baseline_config1_scripted(gpu, oracle, seed):
    """BASELINE.json configs[0]: one env, data/config-mini.json, 200 scripted keys (all 19 keys of
    KeyMap::ai incl. the MoveUntil capitals, '>' and search) - seed 0 (rand_xorshift's zero-seed
    substitution) and the file's own seed 4."""
    with open(os.path.join(ROOT, "tests", "golden", "config_mini.json")) as f:
        cfg = json.load(f)
    cfg["seed"] = seed
    script = (SCRIPT * 4)[:200]
    assert set(script) >= set(".hjklnbuy>sHJKLNBUY") and len(script) == 200
    keys = np.frombuffer(script.encode(), np.uint8)
    _lockstep(gpu, oracle, cfg, 1, [seed], 1000, lambda t: keys[t:t + 1], 200, dump_every=10)
    # the same through the single-env class the gym wrapper drives
    g = gpu.GameState(1000, json.dumps(cfg))
    o = oracle.OracleEnv(cfg, max_steps=1000)
    for k in keys:
        g.react(int(k))
        o.react(int(k))
    assert g.prev().dungeon == o.dungeon()
    assert len(json.loads(g.dump_history())) == 200


@pytest.mark.parametrize("mode", ["lines", "direct"])
def test_host_mirror_at_full_size(gpu, cabi, oracle, monkeypatch, mode):
    """The host mirror at 65 536 envs: ~50 k changed lines per step, and the simultaneous end of every episode at step
    60 (every screen is redrawn). The mirror must equal a full copy of the block."""
    monkeypatch.setenv("RG_MIRROR_MODE", mode)
    n, max_steps = 65536, 60
    seeds = np.arange(n, dtype=np.uint64) + np.uint64(1)
    L, h = _raw_batch(cabi, {}, n, max_steps, seeds)
    obs, hist = cabi.HostObs(), C.c_void_p()
    rc = L.rg_mirror_get(h, C.byref(obs), C.byref(hist))
    assert rc in (0, 3), L.rg_last_error(h)

    def view(ptr, ctype, count):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,))

    v = cabi.Views()
    cabi.check(L.rg_views_get(h, C.byref(v)), h)
    m_screen, m_status = view(obs.screen, C.c_uint8, n * 1920), view(obs.status, C.c_uint32, n * 10)
    m_reward, m_done = view(obs.reward, C.c_int32, n), view(obs.done, C.c_uint8, n)
    m_message, m_error = view(obs.message, C.c_uint32, n), view(obs.error, C.c_uint8, n)
    m_hist = view(hist, C.c_uint8, n * v.hist_stride).reshape(n, v.hist_stride)
    screen, history = np.zeros((n, 1920), np.uint8), np.zeros((n, 1920), np.uint8)
    status, reward = np.zeros((n, 10), np.uint32), np.zeros(n, np.int32)
    done, message, error = np.zeros(n, np.uint8), np.zeros(n, np.uint32), np.zeros(n, np.uint8)
    full = cabi.HostObs(screen.ctypes.data, history.ctypes.data, status.ctypes.data, reward.ctypes.data, done.ctypes.data,
                        message.ctypes.data, error.ctypes.data)
    ids = np.arange(n, dtype=np.uint64)
    nbytes = C.c_uint64()
    sent = []
    for t in range(75):
        keys = oracle.synthetic_actions(t, ids)
        rc = L.rg_step_mirror(h, keys.ctypes.data, 1, C.byref(nbytes))
        assert rc in (0, 3), L.rg_last_error(h)
        sent.append(nbytes.value)
        if t % 10 == 9 or t in (58, 59, 60, 61):
            L.rg_fetch(h, C.byref(full))
            assert np.array_equal(m_screen.reshape(n, -1), screen), "screen differs at step %d" % t
            bits = np.unpackbits(m_hist, axis=1, bitorder="little")[:, :1920]
            assert np.array_equal(bits, history), "visited map differs at step %d" % t
            assert np.array_equal(m_status.reshape(n, 10), status) and np.array_equal(m_reward, reward), t
            assert np.array_equal(m_done, done) and np.array_equal(m_message, message) and np.array_equal(m_error, error), t
    assert max(sent) > 8e6 and np.median(sent) < 8e6, (max(sent), np.median(sent))  # the mass reset; a normal step
    L.rg_destroy(h)
