"""Adapters for the `rainy` RL library, as in the reference (python/rogue_gym/rainy_impls.py).
Importable only when `rainy` is installed. `ParallelRogueEnvExt.extract` encodes the whole batch
with ONE device launch (`rg_encode_states`) instead of a Python loop over `ImageSetting.expand`
+ np.stack (rainy_impls.py:65-66); for a device-resident pipeline use `rogue_gym.envs.DeviceRogueEnv`."""
from typing import Iterable, Tuple

import numpy as np
from numpy import ndarray

try:
    from rainy.envs import EnvExt, EnvSpec, ParallelEnv
    from rainy.prelude import Array
except ImportError:
    raise ImportError("To use rogue_gym.rainy_impls, install rainy first.")

from ._gymapi import Env
from .envs.parallel import ParallelRogueEnv
from .envs.rogue_env import PlayerState, RogueEnv
from .envs.wrappers import check_rogue_env

ACTION_DIM = len(RogueEnv.ACTIONS)


class RogueEnvExt(EnvExt):
    def __init__(self, env: Env) -> None:
        check_rogue_env(env)
        super().__init__(env)

    @property
    def action_dim(self) -> int:
        return ACTION_DIM

    @property
    def state_dim(self) -> Tuple[int, ...]:
        return self._env.unwrapped.observation_space.shape

    def extract(self, state: PlayerState) -> ndarray:
        return self._env.unwrapped.image_setting.expand(state)

    def save_history(self, file_name: str) -> None:
        self._env.unwrapped.save_actions(file_name)


class ParallelRogueEnvExt(ParallelEnv):
    def __init__(self, env: ParallelRogueEnv) -> None:
        self._env = env
        self._spec = EnvSpec(env.observation_space.shape, env.action_space)

    def close(self) -> None:
        self._env.close()

    def reset(self) -> "Array[PlayerState]":
        return np.array(self._env.reset())

    def step(self, actions: Iterable[int]):
        return tuple(map(np.array, self._env.step(actions)))

    def seed(self, seeds: Iterable[int]) -> None:
        self._env.seed(list(seeds))

    @property
    def num_envs(self) -> int:
        return self._env.num_workers

    @property
    def spec(self) -> "EnvSpec":
        return self._spec

    def extract(self, states: Iterable[PlayerState]) -> "Array":
        return self._env.game.encode_states(list(states), *self._env.image_setting.encoder_args())
