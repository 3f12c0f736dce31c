"""Host-side helpers for large rollouts on one or several GPUs: how the env batch is sharded
(contiguous env-id blocks, one process per GPU, no collective on the data path)
and the synthetic action stream of SURVEY.md §8d. No game logic lives here."""
import ctypes as C

import numpy as np

from . import _cabi

KEYS11 = np.frombuffer(b".hjklnbuy>s", np.uint8)  # RogueEnv.ACTIONS, python/rogue_gym/envs/rogue_env.py:159-172


def shard_range(n_total, rank, world_size):
    """Contiguous block [lo, hi) of env ids owned by `rank`; blocks differ by at most one env."""
    if not (0 <= rank < world_size) or n_total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_seeds(env_ids, base_seed=1):
    """Game seed of env i = base_seed + i (base 1 avoids rand_xorshift's zero-seed substitution)."""
    return (np.asarray(env_ids, np.uint64) + np.uint64(base_seed)).astype(np.uint64)


def synthetic_actions(t, env_ids):
    """a[i,t] = splitmix64(0x9E3779B97F4A7C15*(t+1) ^ i) % 11 -> ASCII key; independent of game RNG."""
    with np.errstate(over="ignore"):
        x = (np.uint64(0x9E3779B97F4A7C15) * np.uint64(t + 1)) ^ np.asarray(env_ids, np.uint64)
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return KEYS11[(x % np.uint64(11)).astype(np.int64)]


class Shard:
    """One GPU's block of envs behind the raw C ABI (device-resident stepping, no per-env objects)."""

    def __init__(self, config_json, env_lo, env_hi, max_steps=1000, device=0, base_seed=1, seeds=None):
        """seeds: None = env i gets base_seed + i; else (lo, hi) uint64 arrays, one 128-bit seed per env."""
        self.L = _cabi.lib()
        self.n = env_hi - env_lo
        self.env_ids = np.arange(env_lo, env_hi, dtype=np.uint64)
        arr = (C.c_char_p * 1)(config_json.encode())
        h = C.c_void_p()
        _cabi.check(self.L.rg_create(arr, 1, self.n, max_steps, device, C.byref(h)))
        self.h = h
        self.params = _cabi.Params()
        _cabi.check(self.L.rg_parse_config(config_json.encode(), C.byref(self.params), None, 0))
        self.W, self.H = self.params.width, self.params.height
        self.base_seed = base_seed
        self.seeds = seeds
        self.reseed_and_reset()

    def reseed_and_reset(self):
        if self.seeds is None:
            lo, hi = env_seeds(self.env_ids, self.base_seed), None
        else:
            lo, hi = (np.ascontiguousarray(a, np.uint64) for a in self.seeds)
        _cabi.check(self.L.rg_seed(self.h, lo.ctypes.data, hi.ctypes.data if hi is not None else None), self.h)
        _cabi.check(self.L.rg_reset(self.h), self.h)
        rc = self.L.rg_sync(self.h)
        if rc not in (_cabi.RG_OK, _cabi.RG_ERR_PANIC):
            _cabi.check(rc, self.h)

    def stream(self):
        return self.L.rg_stream(self.h)

    def step_device(self, actions_dev_ptr):
        rc = self.L.rg_step(self.h, actions_dev_ptr, 1)
        if rc != _cabi.RG_OK:
            _cabi.check(rc, self.h)

    def step_host(self, actions_host_ptr, obs):
        """rg_step_host; sticky reference-panic envs are reported by count(), not raised."""
        rc = self.L.rg_step_host(self.h, actions_host_ptr, 1, C.byref(obs))
        if rc not in (_cabi.RG_OK, _cabi.RG_ERR_PANIC):
            _cabi.check(rc, self.h)

    def mirror(self, with_history=False):
        """Host mirror of the observation block (rg_mirror_get): HostObs of pinned pointers. The bit-packed visited
        map is kept current only if asked for."""
        obs, hist = _cabi.HostObs(), C.c_void_p()
        rc = self.L.rg_mirror_get(self.h, C.byref(obs), C.byref(hist) if with_history else None)
        if rc not in (_cabi.RG_OK, _cabi.RG_ERR_PANIC):
            _cabi.check(rc, self.h)
        return obs, hist

    def step_mirror(self, actions_host_ptr):
        """rg_step_mirror; returns the bytes stored into the host mirror by this step."""
        nbytes = C.c_uint64()
        rc = self.L.rg_step_mirror(self.h, actions_host_ptr, 1, C.byref(nbytes))
        if rc not in (_cabi.RG_OK, _cabi.RG_ERR_PANIC):
            _cabi.check(rc, self.h)
        return nbytes.value

    def quiesce(self):
        """Main stream waits for the background next-episode generation queued so far (async)."""
        _cabi.check(self.L.rg_quiesce(self.h), self.h)

    def sync(self):
        rc = self.L.rg_sync(self.h)
        if rc not in (_cabi.RG_OK, _cabi.RG_ERR_PANIC):
            _cabi.check(rc, self.h)

    def errors(self):
        err = np.zeros(self.n, np.uint8)
        obs = _cabi.HostObs(None, None, None, None, None, None, err.ctypes.data)
        _cabi.check(self.L.rg_fetch(self.h, C.byref(obs)), self.h)
        return err

    def hashes(self):
        out = np.zeros(self.n, np.uint64)
        _cabi.check(self.L.rg_state_hash(self.h, out.ctypes.data), self.h)
        return out

    def stats(self):
        out = np.zeros(8, np.uint64)
        _cabi.check(self.L.rg_stats(self.h, out.ctypes.data), self.h)
        names = ("swap_in", "sync_reset", "full_step", "prefetch_built", "prefetch_stale", "monster_env_steps", "bfs_levels",
                 "fast_steps")
        return {k: int(v) for k, v in zip(names, out)}

    def launches(self):
        return int(self.L.rg_launch_count(self.h))

    def close(self):
        if self.h:
            self.L.rg_destroy(self.h)
            self.h = None
