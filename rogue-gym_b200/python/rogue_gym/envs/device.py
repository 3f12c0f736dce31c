"""DeviceRogueEnv — the batched, device-resident face of the simulator for trainers.

The reference hands a trainer one Python `PlayerState` per env and the trainer expands each into a
float32 image and stacks them (the reference's python/rogue_gym/rainy_impls.py:65-66: a Python loop + np.stack, the
401 KB/env-step cost of the default observation crosses PCIe and the interpreter). Here nothing
leaves HBM: actions come in as a CUDA tensor, the step kernels run on the batch's stream, the image
encoder (`rg_encode`, same planes as `ImageSetting.expand`) writes straight into a torch tensor the
policy reads, and screen / status / reward / done are zero-copy tensor views of the C ABI's
`rg_views` block. No per-env Python objects are created.

    env = DeviceRogueEnv({"seed": None}, num_envs=65536, image_setting=ImageSetting(), seeds=range(1, 65537))
    obs = env.reset()                                  # CompactObs(symbols u8 [N,H,W], status i32 [N,9], history) on cuda
    obs, reward, done, info = env.step(actions)        # actions: int tensor [N], indices into ACTIONS
    image = env.expand(obs)                            # float32 [N, C, H, W] == ImageSetting.expand per state
(`observation="image"` makes every step write that image instead: 401 KB per env for the reference default.)

Semantics are those of `ParallelRogueEnv` (auto-reset: a terminal env returns the first observation
of its next episode with done = 1; reward = max(0, gold gained), parallel.py:60-66) plus the optional
stair bonus of `StairRewardParallel` (wrappers.py:44-64). Envs in a state where the reference's worker
thread would have panicked stay frozen and are reported by `errors()`; nothing is raised per step.
"""
import ctypes as C
import json
from typing import Iterable, NamedTuple, Optional, Sequence

import numpy as np

from rogue_gym_python import _cabi

from .._gymapi import spaces
from .rogue_env import DungeonType, ImageSetting, RogueEnv


class CompactObs(NamedTuple):
    """The default observation of DeviceRogueEnv: what `ImageSetting.expand` carries, before the one-hot expansion.
    symbols uint8 [N, H, W] (Symbol ids, core/src/symbol.rs:17-40), status int32 [N, 9] (StatusFlag order),
    history uint8 [N, H, W] 0/1 or None. The tensors are the env's own buffers, overwritten by the next step."""

    symbols: "object"
    status: "object"
    history: "object"


class SymbolExpand:
    """CompactObs -> the float32 [N, C, H, W] image of `ImageSetting.expand` (python/rogue_gym/envs/rogue_env.py:84-98,
    python/src/lib.rs:158-205), on the device, bit for bit: one-hot symbol planes (the last one is always empty,
    SURVEY.md 8a a14), one constant plane per selected status value, the visited map. A policy would rather feed
    `obs.symbols.long()` to an nn.Embedding and `obs.status` to its head; this is the proof that nothing is lost."""

    def __init__(self, image_setting: "ImageSetting", symbols: int):
        self.setting, self.symbols = image_setting, int(symbols)

    def __call__(self, obs: CompactObs):
        import torch
        import torch.nn.functional as F
        n, h, w = obs.symbols.shape
        planes = []
        if self.setting.dungeon == DungeonType.SYMBOL:
            hot = F.one_hot(obs.symbols.long(), self.symbols).permute(0, 3, 1, 2).to(torch.float32)
            hot[:, self.symbols - 1] = 0
            planes.append(hot)
        else:
            # one IEEE f32 division per cell like python/src/lib.rs:72-84 (a Python-scalar divisor would be turned
            # into a multiplication by its reciprocal, which rounds differently)
            div = torch.full((), float(self.symbols), dtype=torch.float32, device=obs.symbols.device)
            planes.append((obs.symbols.to(torch.float32) / div).unsqueeze(1))
        flag = self.setting.status.value
        cols = [k for k in range(9) if flag & (1 << k)]
        if cols:
            planes.append(obs.status[:, cols].to(torch.float32)[:, :, None, None].expand(n, len(cols), h, w))
        if self.setting.includes_hist:
            planes.append(obs.history.to(torch.float32).unsqueeze(1))
        return torch.cat(planes, dim=1)


class _DevPtr:
    """Minimal __cuda_array_interface__ carrier for a pointer the C ABI owns."""

    def __init__(self, ptr, shape, typestr, strides=None):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": strides}


class DeviceRogueEnv:
    metadata = RogueEnv.metadata
    SYMBOLS = RogueEnv.SYMBOLS
    ACTION_MEANINGS = RogueEnv.ACTION_MEANINGS
    ACTIONS = RogueEnv.ACTIONS
    ACTION_LEN = len(ACTIONS)

    def __init__(
        self,
        config_dict: Optional[dict] = None,
        num_envs: int = 1,
        max_steps: int = 1000,
        image_setting: ImageSetting = ImageSetting(),
        device: int = 0,
        seeds: Optional[Iterable[int]] = None,
        stair_reward: float = 0.0,
        observation: str = "compact",
    ) -> None:
        """observation: "compact" (default) - every step returns a CompactObs of device tensors (symbol ids + status
        vector [+ visited map if image_setting.includes_hist]: 1.9 KB per env, `env.expand(obs)` gives the image);
        "image" - the float32 [N, C, H, W] image of `image_setting` (401 KB per env for the reference default)."""
        if observation not in ("compact", "image"):
            raise ValueError("observation must be 'compact' or 'image'")
        self.observation = observation
        import torch

        self._torch = torch
        self.num_envs = self.num_workers = int(num_envs)
        self.max_steps = max_steps
        self.image_setting = image_setting
        self.stair_reward = float(stair_reward)
        self.device = torch.device("cuda", device)
        self._L = _cabi.lib()
        cfg = json.dumps(config_dict or {})
        arr = (C.c_char_p * 1)(cfg.encode())
        h = C.c_void_p()
        _cabi.check(self._L.rg_create(arr, 1, self.num_envs, int(max_steps), device, C.byref(h)))
        self._h = h
        v = _cabi.Views()
        _cabi.check(self._L.rg_views_get(h, C.byref(v)), h)
        n, self.height, self.width = self.num_envs, v.height, v.width
        as_t = lambda ptr, shape, ts, strides=None: torch.as_tensor(_DevPtr(ptr, shape, ts, strides), device=self.device)
        # zero-copy views of the observation block; refreshed in place by every step / reset
        self.screen = as_t(v.screen, (n, v.height, v.width), "|u1", (v.cell_stride, v.width, 1))
        self.status = as_t(v.status, (n, 10), "<i4")
        self.gold_reward = as_t(v.reward, (n,), "<i4")
        self.done = as_t(v.done, (n,), "|u1")
        self.message = as_t(v.message, (n,), "<i4")
        self.error = as_t(v.error, (n,), "|u1")
        params = _cabi.Params()
        _cabi.check(self._L.rg_parse_config(cfg.encode(), C.byref(params), None, 0))
        self.symbols = int(params.symbols)
        self.action_space = spaces.discrete.Discrete(self.ACTION_LEN)
        self.observation_space = image_setting.detect_space(v.height, v.width, self.symbols)
        self._enc = image_setting.encoder_args()
        self.channels = int(self._L.rg_encode_channels(h, *self._enc))
        self.expand = SymbolExpand(image_setting, self.symbols)
        if observation == "image":
            self.obs = torch.empty((n, self.channels, v.height, v.width), dtype=torch.float32, device=self.device)
        else:
            hist = torch.empty((n, v.height, v.width), dtype=torch.uint8, device=self.device) if image_setting.includes_hist else None
            self.obs = CompactObs(torch.empty((n, v.height, v.width), dtype=torch.uint8, device=self.device),
                                  torch.empty((n, 9), dtype=torch.int32, device=self.device), hist)
        self.reward = torch.zeros(n, dtype=torch.float32, device=self.device)
        self._stream = torch.cuda.ExternalStream(int(self._L.rg_stream(h)), device=self.device)
        if seeds is not None:
            self.seed(seeds)
            self.reset()
        else:
            self._observe()

    # ---- plumbing: everything the env launches goes to the batch's own stream, ordered after the
    # caller's current stream on the way in and before it on the way out
    def _enter(self):
        cur = self._torch.cuda.current_stream(self.device)
        self._stream.wait_stream(cur)
        return cur

    def _observe(self):
        ch = C.c_int()
        with self._torch.cuda.stream(self._stream):
            if self.observation == "image":
                _cabi.check(self._L.rg_encode(self._h, self._enc[0], self._enc[1], self._enc[2], self.obs.data_ptr(),
                                              C.byref(ch)), self._h)
            else:
                o = self.obs
                _cabi.check(self._L.rg_encode_compact(self._h, o.symbols.data_ptr(), o.status.data_ptr(),
                                                      o.history.data_ptr() if o.history is not None else None), self._h)

    def _alive(self):
        if not self._h:
            raise RuntimeError("Error in rogue-gym: the game state was closed")

    def seed(self, seeds: Sequence[int]) -> None:
        """One seed per env (up to 128 bits); takes effect at the next reset."""
        self._alive()
        seeds = [int(s) for s in seeds]
        if len(seeds) != self.num_envs:
            raise ValueError("expected %d seeds, got %d" % (self.num_envs, len(seeds)))
        lo = np.array([s & 0xFFFFFFFFFFFFFFFF for s in seeds], np.uint64)
        hi = np.array([(s >> 64) & 0xFFFFFFFFFFFFFFFF for s in seeds], np.uint64)
        _cabi.check(self._L.rg_seed(self._h, lo.ctypes.data, hi.ctypes.data), self._h)

    def reset(self):
        self._alive()
        cur = self._enter()
        _cabi.check(self._L.rg_reset(self._h), self._h)
        _cabi.check(self._L.rg_train_reset(self._h), self._h)
        self._observe()
        with self._torch.cuda.stream(self._stream):
            self.reward.zero_()
        cur.wait_stream(self._stream)
        return self.obs

    def step(self, actions):
        """actions: integer tensor / array [N] of indices into ACTIONS (uint8, int32 or int64 on the device
        are used as they are). One C-ABI call (`rg_step_train`): index -> key, the step kernels, the image
        encoder into `obs`, the float reward with the stair bonus - all on the batch's stream. An index
        outside 0..10 leaves that env untouched and marks it in `error` (the reference raises ValueError)."""
        torch = self._torch
        self._alive()
        a = torch.as_tensor(actions, device=self.device)
        if a.shape != (self.num_envs,):
            raise ValueError("Invalid action: expected shape (%d,), got %s" % (self.num_envs, tuple(a.shape)))
        if a.dtype not in (torch.uint8, torch.int32, torch.int64):
            a = a.long()
        return self._step(a.contiguous(), a.element_size())

    def step_keys(self, keys):
        """keys: uint8 tensor / array [N] of ASCII keys (capitals = move until blocked)."""
        torch = self._torch
        self._alive()
        k = torch.as_tensor(keys, device=self.device).to(torch.uint8).contiguous()
        if k.shape != (self.num_envs,):
            raise ValueError("Invalid action: expected shape (%d,), got %s" % (self.num_envs, tuple(k.shape)))
        return self._step(k, 0)

    def _step(self, a, index_bytes):
        cur = self._enter()
        # the kernels read `a` on the batch's stream: keep it alive until the next step instead of
        # record_stream() (the batch owns that stream and destroys it in close(), which the caching
        # allocator's bookkeeping for recorded streams does not survive at interpreter exit)
        self._last_actions = a
        image = self.observation == "image"
        _cabi.check(self._L.rg_step_train(self._h, a.data_ptr(), index_bytes, self._enc[0], self._enc[1],
                                          self._enc[2], self.obs.data_ptr() if image else None, self.reward.data_ptr(),
                                          C.c_float(self.stair_reward)), self._h)
        if not image:
            self._observe()
        cur.wait_stream(self._stream)
        return self.obs, self.reward, self.done, {}

    def history(self):
        """uint8 [N, H, W] visited map (PlayerState.history), unpacked on the device."""
        torch = self._torch
        v = _cabi.Views()
        _cabi.check(self._L.rg_views_get(self._h, C.byref(v)), self._h)
        bits = torch.as_tensor(_DevPtr(v.history_bits, (self.num_envs, v.hist_stride), "|u1"), device=self.device)
        cells = self.height * self.width
        idx = torch.arange(cells, device=self.device)
        return ((bits[:, idx >> 3] >> (idx & 7).to(torch.uint8)) & 1).reshape(self.num_envs, self.height, self.width)

    def errors(self) -> np.ndarray:
        """uint8 [N] rg_status per env (0 = fine, 3 = frozen in a state where the reference panics)."""
        self._torch.cuda.current_stream(self.device).synchronize()
        return self.error.cpu().numpy()

    def close(self) -> None:
        self._last_actions = None
        if self._h:
            self._torch.cuda.synchronize(self.device)
            self._L.rg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
