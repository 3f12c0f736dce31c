"""ParallelRogueEnv — N games stepped in lockstep with auto-reset, the reference's vector env
(python/rogue_gym/envs/parallel.py:8-77) on top of `ParallelGameState` of the B200 module.

The list-of-PlayerState interface is kept for drop-in use (the reference's tests and the rainy
adapter consume it); for large batches use `rogue_gym.envs.DeviceRogueEnv`, which keeps
observations in HBM and never builds per-env Python objects.
Differences from the reference: `get_key_to_action` works (attribute typo in the reference,
parallel.py:37-38) and `get_configs` returns one dict per worker (the reference calls a
`dump_config` its ParallelGameState does not have, parallel.py:40-42)."""
import json
from typing import Dict, Iterable, List, Tuple, Union

from rogue_gym_python._config import dump_config
from rogue_gym_python._rogue_gym import ParallelGameState, PlayerState

from .._gymapi import spaces
from .rogue_env import ImageSetting, RogueEnv


class ParallelRogueEnv:
    metadata = RogueEnv.metadata
    SYMBOLS = RogueEnv.SYMBOLS
    ACTION_MEANINGS = RogueEnv.ACTION_MEANINGS
    ACTIONS = RogueEnv.ACTIONS
    ACTION_LEN = len(ACTIONS)

    def __init__(
        self,
        config_dicts: Iterable[dict],
        max_steps: int = 1000,
        image_setting: ImageSetting = ImageSetting(),
    ) -> None:
        self._configs = [json.dumps(d) for d in config_dicts]
        self.game = ParallelGameState(max_steps, self._configs)
        self.result = None
        self.max_steps = max_steps
        self.steps = 0
        self.action_space = spaces.discrete.Discrete(self.ACTION_LEN)
        self.observation_space = image_setting.detect_space(*self.game.screen_size(), self.game.symbols())
        self.image_setting = image_setting
        self.states = self.game.states()
        self.num_workers = len(self._configs)

    def get_key_to_action(self) -> Dict[str, str]:
        return self.ACTION_MEANINGS

    def get_configs(self) -> List[dict]:
        return [json.loads(dump_config(c)) for c in self._configs]

    def _keys(self, action: Union[Iterable[int], str]) -> List[int]:
        # a string with one key per worker is taken literally (this is how capitals = MoveUntil are
        # reachable); anything else is a sequence of indices into ACTIONS (parallel.py:52-58)
        if isinstance(action, str) and len(action) == self.num_workers:
            return [ord(c) for c in action]
        try:
            return [ord(self.ACTIONS[x]) for x in action]
        except Exception:
            raise ValueError("Invalid action: {}".format(action))

    def step(
        self, action: Union[Iterable[int], str]
    ) -> Tuple[List[PlayerState], List[float], List[bool], List[dict]]:
        states = self.game.step(self._keys(action))
        rewards = [max(0, new.gold - old.gold) for old, new in zip(self.states, states)]
        self.states = states
        return states, rewards, [s.is_terminal for s in states], [{}] * self.num_workers

    def reset(self) -> List[PlayerState]:
        self.states = self.game.reset()
        return self.states

    def close(self) -> None:
        self.game.close()

    def seed(self, seeds: List[int]) -> None:
        self.game.seed(seeds)
