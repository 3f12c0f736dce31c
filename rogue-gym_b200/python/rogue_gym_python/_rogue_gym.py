"""`rogue_gym_python._rogue_gym` — same classes, method names, argument meaning and error
behaviour as the reference's PyO3 module (python/src/lib.rs:208-366), backed by the CUDA
library through include/rogue_b200.h.

    GameState(max_steps, config_str)          python/src/lib.rs:208-258
    ParallelGameState(max_steps, configs)     python/src/lib.rs:260-335
    PlayerState                               python/src/lib.rs:29-206

plus, beyond the reference surface, `ParallelGameState.step_arrays` / `.views()` for batched
consumers that do not want one Python object per env.
"""
import ctypes as C
import json

import numpy as np

from . import _cabi
from ._cabi import check

STATUS_KEYS = ("dungeon_level", "gold", "hp_current", "hp_max", "str_current", "str_max", "defense",
               "player_level", "exp", "hunger")  # Status::to_dict_vec core/src/character/player.rs:403-416
_HUNGER = ("", "Hungry", "Weak")  # Display of Hunger, player.rs:362-386 (Normal prints nothing)
_FLAG_ORDER = (0, 2, 3, 4, 5, 6, 7, 8, 9)  # StatusFlagInner::to_vector python/src/flags.rs:63-85

# KeyMap::ai (core/src/input.rs:74-99) as the JSON the reference's dump_history emits
_DIRS = {"l": "Right", "k": "Up", "j": "Down", "h": "Left", "u": "RightUp", "y": "LeftUp", "n": "RightDown",
         "b": "LeftDown"}


def _input_code(key):
    ch = chr(key)
    if ch in _DIRS:
        return {"Act": {"Move": _DIRS[ch]}}
    if ch.lower() in _DIRS and ch.isupper():
        return {"Act": {"MoveUntil": _DIRS[ch.lower()]}}
    return {"Act": {"s": "Search", ".": "NoOp", ">": "DownStair"}[ch]}


_KEY_OF_DIR = {v: k for k, v in _DIRS.items()}


def keys_from_history(text):
    """Inverse of `dump_history`: the JSON list RunTime::saved_inputs_as_json writes (core/src/lib.rs:
    357-375; entries like {"Act": {"Move": "RightUp"}} or {"Act": "Search"}, InputCode / Action of
    core/src/input.rs:24-60) back to the ASCII keys of KeyMap::ai, so that a recorded episode - e.g.
    the reference's data/learned/*/best-actions.json - can be re-simulated."""
    entries = json.loads(text) if isinstance(text, (str, bytes)) else text
    keys = bytearray()
    for e in entries:
        act = e.get("Act") if isinstance(e, dict) else None
        if isinstance(act, dict) and len(act) == 1:
            (kind, direction), = act.items()
            key = _KEY_OF_DIR.get(direction)
            if key is not None and kind == "Move":
                keys.append(ord(key))
                continue
            if key is not None and kind == "MoveUntil":
                keys.append(ord(key.upper()))
                continue
        elif act in ("Search", "NoOp", "DownStair"):
            keys.append(ord({"Search": "s", "NoOp": ".", "DownStair": ">"}[act]))
            continue
        raise ValueError("unsupported entry in action history: %r (only what KeyMap::ai can produce is replayable)" % (e,))
    return bytes(keys)


class _Batch:
    """Owns one rg_batch handle."""

    def __init__(self, configs, n_envs, max_steps, device=0):
        self.L = _cabi.lib()
        arr = (C.c_char_p * len(configs))(*[c.encode() for c in configs])
        h = C.c_void_p()
        rc = self.L.rg_create(arr, len(configs), n_envs, max_steps, device, C.byref(h))
        if rc == _cabi.RG_ERR_PARSE:
            raise RuntimeError(_cabi.last_error())  # already "Failed to parse config: ..."
        if rc == _cabi.RG_ERR_SETTING:
            msg = _cabi.last_error()
            raise RuntimeError(msg if msg.startswith("Error in rogue-gym") else "Error in rogue-gym: " + msg)
        check(rc)
        self.h = h
        self.n = n_envs
        self.params = _cabi.Params()
        check(self.L.rg_parse_config(configs[0].encode(), C.byref(self.params), None, 0))
        self.W, self.H = self.params.width, self.params.height
        self.C = self.W * self.H
        n = n_envs
        self.screen = np.zeros((n, self.C), np.uint8)
        self.history = np.zeros((n, self.C), np.uint8)
        self.status = np.zeros((n, 10), np.uint32)
        self.reward = np.zeros(n, np.int32)
        self.done = np.zeros(n, np.uint8)
        self.message = np.zeros(n, np.uint32)
        self.error = np.zeros(n, np.uint8)
        self.obs = _cabi.HostObs(self.screen.ctypes.data, self.history.ctypes.data, self.status.ctypes.data,
                                 self.reward.ctypes.data, self.done.ctypes.data, self.message.ctypes.data,
                                 self.error.ctypes.data)

    def close(self):
        if getattr(self, "h", None):
            self.L.rg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _alive(self):
        if not self.h:
            raise RuntimeError("Error in rogue-gym: the game state was closed")

    def seed(self, seeds):
        self._alive()
        lo = np.array([s & 0xFFFFFFFFFFFFFFFF for s in seeds], np.uint64)
        hi = np.array([(s >> 64) & 0xFFFFFFFFFFFFFFFF for s in seeds], np.uint64)
        check(self.L.rg_seed_first(self.h, lo.ctypes.data, hi.ctypes.data, len(seeds)), self.h)

    def reset(self):
        self._alive()
        check(self.L.rg_reset(self.h), self.h)
        check(self.L.rg_sync(self.h), self.h)
        self.fetch()

    def fetch(self):
        self._alive()
        check(self.L.rg_fetch(self.h, C.byref(self.obs)), self.h)

    def step(self, keys, auto_reset):
        self._alive()
        a = np.ascontiguousarray(keys, dtype=np.uint8)
        if a.shape != (self.n,):
            raise RuntimeError("Error in rogue-gym: expected %d actions, got %s" % (self.n, a.shape))
        check(self.L.rg_step_host(self.h, a.ctypes.data, int(auto_reset), C.byref(self.obs)), self.h)

    # ---- host mirror (include/rogue_b200.h rg_mirror_*): numpy views of pinned memory the device
    # keeps current by writing only what changed; read-only for the caller
    def mirror(self):
        self._alive()
        if getattr(self, "_mirror", None) is None:
            obs, hist = _cabi.HostObs(), C.c_void_p()
            check(self.L.rg_mirror_get(self.h, C.byref(obs), C.byref(hist)), self.h, tolerate_env_errors=True)
            v = _cabi.Views()
            check(self.L.rg_views_get(self.h, C.byref(v)), self.h)
            n = self.n

            def view(ptr, ctype, shape):
                count = int(np.prod(shape))
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).reshape(shape)

            self._mirror = dict(
                screen=view(obs.screen, C.c_uint8, (n, self.H, self.W)),
                history_bits=view(hist, C.c_uint8, (n, v.hist_stride)),
                status=view(obs.status, C.c_uint32, (n, 10)), reward=view(obs.reward, C.c_int32, (n,)),
                done=view(obs.done, C.c_uint8, (n,)), message=view(obs.message, C.c_uint32, (n,)),
                error=view(obs.error, C.c_uint8, (n,)))
        return self._mirror

    def mirror_sync(self):
        nbytes = C.c_uint64()
        check(self.L.rg_mirror_sync(self.h, C.byref(nbytes)), self.h, tolerate_env_errors=True)
        return nbytes.value

    def step_mirror(self, keys, auto_reset=True):
        """One step whose observation lands in mirror(); returns the bytes that crossed PCIe for it."""
        m = self.mirror()
        a = np.ascontiguousarray(keys, dtype=np.uint8)
        if a.shape != (self.n,):
            raise RuntimeError("Error in rogue-gym: expected %d actions, got %s" % (self.n, a.shape))
        nbytes = C.c_uint64()
        check(self.L.rg_step_mirror(self.h, a.ctypes.data, int(auto_reset), C.byref(nbytes)), self.h,
              tolerate_env_errors=True)
        del m
        return nbytes.value

    def mirror_history(self):
        """uint8 [N, H, W] 0/1 visited map from the mirrored bit rows."""
        m = self.mirror()
        bits = np.unpackbits(m["history_bits"], axis=1, bitorder="little")[:, :self.C]
        return bits.reshape(self.n, self.H, self.W)

    def player_state(self, i, terminal=None):
        """terminal: None = what the last step returned (`done`); else the state's own flag (rg_fetch_terminal)."""
        return PlayerState(self, self.screen[i].copy(), self.history[i].copy(), self.status[i].copy(),
                           int(self.message[i]), bool(self.done[i] if terminal is None else terminal[i]))

    def state_terminal(self):
        self._alive()
        out = np.zeros(self.n, np.uint8)
        check(self.L.rg_fetch_terminal(self.h, out.ctypes.data), self.h)
        return out

    def set_panic_policy(self, policy):
        self._alive()
        check(self.L.rg_set_panic_policy(self.h, {"sticky": 0, "terminal": 1}.get(policy, policy)), self.h)


class PlayerState:
    """A memory efficient representation of Agent observation (python/src/lib.rs:29-38): a value."""

    def __init__(self, batch, screen, history, status, message, is_terminal):
        self._batch = batch
        self._screen = screen      # uint8 [H*W]
        self._history = history    # uint8 [H*W] 0/1
        self._status = status      # uint32 [10], Status::to_vec order
        self._message = message
        self._is_terminal = is_terminal

    # -- value semantics (#[derive(Clone, PartialEq)])
    def __eq__(self, other):
        if not isinstance(other, PlayerState):
            return NotImplemented
        return (np.array_equal(self._screen, other._screen) and np.array_equal(self._history, other._history)
                and np.array_equal(self._status, other._status) and self.symbols == other.symbols
                and self._message == other._message and self._is_terminal == other._is_terminal)

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = None

    def __repr__(self):  # python/src/lib.rs:116-124 + Status Display player.rs:433-450
        s = self._status
        lines = "".join(row + "\n" for row in self.dungeon)
        return lines + "Level: %2d Gold: %5d Hp: %2d(%2d) Str: %2d(%2d) Arm: %2d Exp: %2d/%2d %s" % (
            s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], _HUNGER[min(int(s[9]), 2)])

    __str__ = __repr__

    @property
    def status(self):
        return {k: int(v) for k, v in zip(STATUS_KEYS, self._status)}

    @property
    def dungeon(self):
        W = self._batch.W
        raw = self._screen.tobytes()
        return [raw[i:i + W].decode("latin-1") for i in range(0, len(raw), W)]

    @property
    def dungeon_level(self):
        return int(self._status[0])

    @property
    def gold(self):
        return int(self._status[1])

    @property
    def symbols(self):
        return int(self._batch.params.symbols)

    @property
    def is_terminal(self):
        return self._is_terminal

    def status_vec(self, flag):
        return [int(np.int32(self._status[_FLAG_ORDER[b]])) for b in range(9) if flag & (1 << b)]

    def _encode(self, mode, flag, with_hist):
        b = self._batch
        b._alive()
        flag = 0 if flag is None else int(flag)
        ch = b.L.rg_encode_channels(b.h, mode, flag, int(with_hist))
        out = np.empty((ch, b.H, b.W), np.float32)
        scr = np.ascontiguousarray(self._screen)
        hist = np.ascontiguousarray(self._history)
        st = np.ascontiguousarray(self._status)
        check(b.L.rg_encode_states(b.h, 1, scr.ctypes.data, hist.ctypes.data, st.ctypes.data, mode, flag,
                                   int(with_hist), out.ctypes.data, None), b.h)
        return out

    def gray_image(self, flag=None):
        return self._encode(0, flag, False)

    def gray_image_with_hist(self, flag=None):
        return self._encode(0, flag, True)

    def symbol_image(self, flag=None):
        return self._encode(1, flag, False)

    def symbol_image_with_hist(self, flag=None):
        return self._encode(1, flag, True)


def _parse_or_default(config_str):
    if config_str is None:
        return "{}"
    if not isinstance(config_str, str):
        raise TypeError("config must be a JSON string")
    return config_str


class GameState:
    """Single game (python/src/lib.rs:208-258): no auto-reset, errors raise."""

    def __init__(self, max_steps, config_str=None):
        self._config_str = _parse_or_default(config_str)
        self._batch = _Batch([self._config_str], 1, int(max_steps))
        self._batch.fetch()
        self._seed = None
        self._history = []
        self._steps = 0
        self._max_steps = int(max_steps)

    def screen_size(self):
        return (self._batch.H, self._batch.W)

    def set_seed(self, seed):
        seed = int(seed)
        if not 0 <= seed < (1 << 64):
            raise OverflowError("seed must fit in u64")
        self._seed = seed
        self._batch.seed([seed])

    def reset(self):
        self._batch.reset()
        self._history = []
        self._steps = 0

    def prev(self):
        return self._batch.player_state(0)

    def react(self, input):
        key = int(input)
        if not 0 <= key < 256:
            raise OverflowError("input must fit in u8")
        live = self._steps <= self._max_steps  # state_impls.rs:52-54: later calls are no-ops
        try:
            self._batch.step(np.array([key], np.uint8), auto_reset=False)
        except RuntimeError as e:
            # react_to_input records the input before it is refused (core/src/lib.rs:288,314)
            if live and getattr(e, "code", None) == _cabi.RG_ERR_IGNORED_INPUT:
                self._history.append(key)
            raise
        if live:
            self._history.append(key)
            self._steps += 1

    def dump_history(self):
        """RunTime::saved_inputs_as_json (core/src/lib.rs:360-363)."""
        return json.dumps([_input_code(k) for k in self._history], indent=2)

    def dump_config(self):
        from ._config import dump_config
        return dump_config(self._config_str, self._seed)

    def symbols(self):
        return int(self._batch.params.symbols)


class ParallelGameState:
    """N games stepped in lockstep with auto-reset (python/src/lib.rs:260-335, thread_impls.rs)."""

    def __init__(self, max_steps, configs, device=0):
        configs = [_parse_or_default(c) for c in configs]
        if len(configs) == 0:
            raise IndexError("index out of bounds: the len is 0 but the index is 0")  # configs[0] lib.rs:281
        uniform = all(c == configs[0] for c in configs)
        self._batch = _Batch(configs[:1] if uniform else configs, len(configs), int(max_steps), device)
        self._batch.fetch()

    def screen_size(self):
        return (self._batch.H, self._batch.W)

    def symbols(self):
        return int(self._batch.params.symbols)

    def seed(self, seed):
        """ThreadConductor::seed zips the workers with the seeds (thread_impls.rs:45-50): with fewer seeds than
        workers the remaining workers keep the seed they have, extra seeds are ignored."""
        seeds = [int(s) for s in seed]
        self._batch.seed(seeds[: self._batch.n])

    def states(self):
        """Instruction::State (thread_impls.rs:51-60,131): every worker's own state - after an auto-reset step that
        is the fresh game, which is not terminal (only the copy `step` returned was flagged)."""
        self._batch.fetch()
        term = self._batch.state_terminal()
        return [self._batch.player_state(i, term) for i in range(self._batch.n)]

    def step(self, input):
        self._batch.step(np.asarray(list(input) if not isinstance(input, np.ndarray) else input, np.uint8), True)
        return [self._batch.player_state(i) for i in range(self._batch.n)]

    def reset(self):
        self._batch.reset()
        return [self._batch.player_state(i) for i in range(self._batch.n)]

    def close(self):
        self._batch.close()

    def encode_states(self, states, mode, flag, with_hist):
        """float32 [len(states), C, H, W]: the images `PlayerState.{gray,symbol}_image[_with_hist]` give
        one by one (python/src/lib.rs:158-205), produced by one device launch for the whole list."""
        b = self._batch
        b._alive()
        n = len(states)
        ch = b.L.rg_encode_channels(b.h, mode, int(flag), int(with_hist))
        out = np.empty((n, ch, b.H, b.W), np.float32)
        if n == 0:
            return out
        scr = np.ascontiguousarray(np.stack([s._screen for s in states]))
        hist = np.ascontiguousarray(np.stack([s._history for s in states]))
        st = np.ascontiguousarray(np.stack([s._status for s in states]))
        check(b.L.rg_encode_states(b.h, n, scr.ctypes.data, hist.ctypes.data, st.ctypes.data, mode, int(flag),
                                   int(with_hist), out.ctypes.data, None), b.h)
        return out

    # ---- batched surface (not in the reference): arrays instead of per-env objects
    def step_arrays(self, keys):
        """keys: uint8[N] ASCII. Steps every env and returns numpy views of the host mirror: screen
        u8 [N,H,W], history_bits (bit-packed rows, see `_Batch.mirror_history`), status, reward, done,
        message, error. The views are refreshed in place by every call (only what changed crosses
        PCIe); envs in an error state are reported through `error`, nothing is raised."""
        b = self._batch
        b.step_mirror(keys, True)
        return b.mirror()
