"""RogueEnv — the single-game gym environment of the reference (python/rogue_gym/envs/
rogue_env.py:114-296), on top of the B200 extension-module mirror `rogue_gym_python._rogue_gym`.

Same public names, argument meaning and results as the reference class; the game itself runs
in the CUDA library (one env = a batch of one). Differences, all deliberate:
  * `config_dict` is copied before `**kwargs` are merged in (the reference mutates its shared
    default `{}`, so keyword settings leak from one `RogueEnv(...)` to the next,
    rogue_env.py:177,186);
  * `get_key_to_action` returns ACTION_MEANINGS (the reference spells the attribute
    `ACION_MEANINGS` and raises AttributeError, rogue_env.py:208-209);
  * `replay` / `play_cli` raise RuntimeError: the terminal UI is out of scope.
"""
import json
from enum import Enum, Flag
from typing import Dict, List, NamedTuple, Optional, Tuple, Union

import numpy as np
from numpy import ndarray

from rogue_gym_python._rogue_gym import GameState, PlayerState

from .._gymapi import Env, spaces


def _need_state(state) -> None:
    if not isinstance(state, PlayerState):
        raise TypeError("Needs PlayerState, but {} was given".format(type(state)))


class StatusFlag(Flag):
    """Which status entries become constant image planes / vector entries, bit i = i-th entry of
    StatusFlagInner::to_vector (python/src/flags.rs:63-85); reference: rogue_env.py:14-59."""

    EMPTY = 0
    DUNGEON_LEVEL = 1 << 0
    HP_CURRENT = 1 << 1
    HP_MAX = 1 << 2
    STR_CURRENT = 1 << 3
    STR_MAX = 1 << 4
    DEFENSE = 1 << 5
    PLAYER_LEVEL = 1 << 6
    EXP = 1 << 7
    HUNGER = 1 << 8
    FULL = (1 << 9) - 1

    def count_one(self) -> int:
        return bin(self.value & 0x1FF).count("1")

    def symbol_image(self, state: PlayerState) -> ndarray:
        _need_state(state)
        return state.symbol_image(flag=self.value)

    def symbol_image_with_hist(self, state: PlayerState) -> ndarray:
        _need_state(state)
        return state.symbol_image_with_hist(flag=self.value)

    def gray_image(self, state: PlayerState) -> ndarray:
        _need_state(state)
        return state.gray_image(flag=self.value)

    def gray_image_with_hist(self, state: PlayerState) -> ndarray:
        _need_state(state)
        return state.gray_image_with_hist(flag=self.value)

    def status_vec(self, state: PlayerState) -> List[int]:
        _need_state(state)
        return state.status_vec(flag=self.value)


class DungeonType(Enum):
    GRAY = 1
    SYMBOL = 2


class ImageSetting(NamedTuple):
    """How a PlayerState becomes a float32 [channels, H, W] array (rogue_env.py:67-98)."""

    dungeon: DungeonType = DungeonType.SYMBOL
    status: StatusFlag = StatusFlag.FULL
    includes_hist: bool = False

    def dim(self, channels: int) -> int:
        planes = channels if self.dungeon == DungeonType.SYMBOL else 1
        return planes + self.status.count_one() + (1 if self.includes_hist else 0)

    def detect_space(self, h: int, w: int, symbols: int):
        return spaces.box.Box(low=0, high=1, shape=(self.dim(symbols), h, w), dtype=np.float32)

    def expand(self, state: PlayerState) -> ndarray:
        _need_state(state)
        symbol = self.dungeon == DungeonType.SYMBOL
        if self.includes_hist:
            fn = self.status.symbol_image_with_hist if symbol else self.status.gray_image_with_hist
        else:
            fn = self.status.symbol_image if symbol else self.status.gray_image
        return fn(state)

    # the (mode, flag, with_hist) triple of the C ABI's rg_encode
    def encoder_args(self) -> Tuple[int, int, int]:
        return (1 if self.dungeon == DungeonType.SYMBOL else 0, self.status.value, int(self.includes_hist))


# Tile bytes in Symbol order (core/src/symbol.rs:17-40, rogue_env.py:118-162)
_SYMBOLS = list(" @#.-%+^!?])/*:=,") + [chr(c) for c in range(ord("A"), ord("Z") + 1)]
# key -> label, as printed by the reference (its j/k labels are swapped relative to KeyMap::ai,
# core/src/input.rs:74-99: j moves down, k moves up; labels only, kept for compatibility)
_ACTION_MEANINGS = {
    ".": "NO_OPERATION", "h": "MOVE_LEFT", "j": "MOVE_UP", "k": "MOVE_DOWN", "l": "MOVE_RIGHT",
    "n": "MOVE_RIGHTDOWN", "b": "MOVE_LEFTDOWN", "u": "MOVE_RIGHTUP", "y": "MOVE_LEFTUP", ">": "DOWNSTAIR",
    "s": "SEARCH",
}


class RogueEnv(Env):
    metadata = {"render.modes": ["human", "ascii"]}
    SYMBOLS = _SYMBOLS
    ACTION_MEANINGS = _ACTION_MEANINGS
    ACTIONS = list(".hjklnbuy>s")
    ACTION_LEN = len(ACTIONS)

    def __init__(
        self,
        config_path: Optional[str] = None,
        config_dict: Optional[dict] = None,
        max_steps: int = 1000,
        image_setting: ImageSetting = ImageSetting(),
        **kwargs,
    ) -> None:
        super().__init__()
        if config_path:
            with open(config_path, "r") as f:
                config = f.read()
        else:
            merged = dict(config_dict or {})
            merged.update(kwargs)
            config = json.dumps(merged)
        self.game = GameState(max_steps, config)
        self.action_space = spaces.discrete.Discrete(self.ACTION_LEN)
        self.observation_space = image_setting.detect_space(*self.game.screen_size(), self.game.symbols())
        self.image_setting = image_setting
        self.result = self.game.prev()

    def screen_size(self) -> Tuple[int, int]:
        """(height, width)"""
        return self.game.screen_size()

    def get_key_to_action(self) -> Dict[str, str]:
        return self.ACTION_MEANINGS

    def get_dungeon(self) -> List[str]:
        return self.result.dungeon

    def get_config(self) -> dict:
        return json.loads(self.game.dump_config())

    def save_config(self, fname: str) -> None:
        with open(fname, "w") as f:
            f.write(self.game.dump_config())

    def save_actions(self, fname: str) -> None:
        with open(fname, "w") as f:
            f.write(self.game.dump_history())

    def replay_actions(self, fname: str) -> Tuple[PlayerState, float, bool, dict]:
        """Re-simulates an action history written by `save_actions` (or by the reference: the
        `saved_inputs` JSON, core/src/lib.rs:357-375) from a fresh reset. Not in the reference, whose
        `replay` is a terminal viewer; returns what `step` returns for the whole sequence."""
        from rogue_gym_python._rogue_gym import keys_from_history
        with open(fname, "r") as f:
            keys = keys_from_history(f.read())
        self.reset()
        return self.step(keys.decode("ascii"))

    def replay(self, interval_ms: int = 100) -> None:
        raise RuntimeError("Currently replay is only supported on UNIX")  # TUI: out of scope here

    def play_cli(self) -> None:
        raise RuntimeError("CLI playing is only supported on UNIX")  # TUI: out of scope here

    def state_to_image(self, state: PlayerState, setting: Optional[ImageSetting] = None) -> ndarray:
        return (self.image_setting if setting is None else setting).expand(state)

    def _feed(self, keys: str) -> int:
        for key in keys:
            self.game.react(ord(key))
        return len(keys)

    def step(self, action: Union[int, str]) -> Tuple[PlayerState, float, bool, dict]:
        """`action` is an index into ACTIONS or a string of keys, each fed to the game in turn
        (e.g. "hjk", "hh>", capitals = move until blocked)."""
        gold_before = self.result.gold
        if isinstance(action, str):
            self._feed(action)
        else:
            try:
                self._feed(self.ACTIONS[action])
            except Exception as e:
                raise ValueError("Invalid action: {} causes {}".format(action, e))
        self.result = self.game.prev()
        return self.result, self.result.gold - gold_before, self.result.is_terminal, {}

    def seed(self, seed: int) -> None:
        """The seed takes effect at the next reset."""
        self.game.set_seed(seed)

    def render(self, mode: str = "human", close: bool = False) -> None:
        print(self.result)

    def reset(self) -> PlayerState:
        self.game.reset()
        self.result = self.game.prev()
        return self.result

    def __repr__(self):
        return repr(self.result)

    @property
    def unwrapped(self):
        return self
