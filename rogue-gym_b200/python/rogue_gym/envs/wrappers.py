"""Reward-shaping wrappers of the reference (python/rogue_gym/envs/wrappers.py): a bonus whenever
the player reaches a deeper dungeon level than before, and the one-floor task built on it."""
from typing import Iterable, List, Tuple, Union

from .._gymapi import Env, Wrapper
from .parallel import ParallelRogueEnv
from .rogue_env import PlayerState, RogueEnv


def check_rogue_env(env: Env) -> None:
    if not isinstance(env.unwrapped, RogueEnv):
        raise ValueError("env have to be a wrapper of RoguEnv")


class StairRewardEnv(Wrapper):
    """+stair_reward on the step that takes the player below every level seen this episode."""

    def __init__(self, env: Env, stair_reward: float = 50.0) -> None:
        check_rogue_env(env)
        self.stair_reward = stair_reward
        self.current_level = 1
        super().__init__(env)

    def step(self, action: Union[int, str]) -> Tuple[PlayerState, float, bool, dict]:
        state, reward, end, info = self.env.step(action)
        level = self.unwrapped.result.status["dungeon_level"]
        if level > self.current_level:
            self.current_level = level
            reward += self.stair_reward
        return state, reward, end, info

    def reset(self) -> PlayerState:
        self.current_level = 1
        return super().reset()

    def __repr__(self):
        return repr(self.env)


class FirstFloorEnv(StairRewardEnv):
    """The episode ends as soon as the second floor is reached."""

    def step(self, action: Union[int, str]) -> Tuple[PlayerState, float, bool, dict]:
        state, reward, end, info = super().step(action)
        return state, reward, end or self.current_level == 2, info


class StairRewardParallel(ParallelRogueEnv):
    """ParallelRogueEnv with the stair bonus. Like the reference (wrappers.py:57-63) the remembered
    level follows the env, so it falls back to 1 when an env auto-resets."""

    def __init__(self, *args, **kwargs) -> None:
        self.stair_reward = kwargs.pop("stair_reward", 50.0)
        super().__init__(*args, **kwargs)
        self.current_levels = [1] * self.num_workers

    def step(
        self, action: Union[Iterable[int], str]
    ) -> Tuple[List[PlayerState], List[float], List[bool], List[dict]]:
        states, rewards, end, info = super().step(action)
        for i, s in enumerate(states):
            level = s.status["dungeon_level"]
            if level > self.current_levels[i]:
                rewards[i] += self.stair_reward
            self.current_levels[i] = level
        return states, rewards, end, info
