"""Resolves the `gym` API the env classes are written against.

The reference imports OpenAI `gym` (python/rogue_gym/envs/rogue_env.py:3-4, wrappers.py:3). This
image has neither `gym` nor `gymnasium`, so when both are missing a minimal stand-in is built
here and registered under the name `gym` (modules `gym`, `gym.spaces`, `gym.spaces.discrete`,
`gym.spaces.box`), which is all the reference's env layer and its tests touch: `gym.Env`,
`gym.Wrapper`, `spaces.discrete.Discrete`, `spaces.box.Box` with value equality. A real `gym`
(or `gymnasium`) always wins when it is importable.
"""
import sys
import types

import numpy as np


def _build_shim():
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    discrete = types.ModuleType("gym.spaces.discrete")
    box = types.ModuleType("gym.spaces.box")

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def __ne__(self, other):
            return not self == other

    class Discrete(Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def sample(self):
            return int(self._rng.randint(self.n))

        def contains(self, x):
            return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

        def __eq__(self, other):
            return isinstance(other, Discrete) and self.n == other.n

        def __repr__(self):
            return "Discrete(%d)" % self.n

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low)
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(np.asarray(low, self.dtype), self.shape)
            self.high = np.broadcast_to(np.asarray(high, self.dtype), self.shape)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

        def __eq__(self, other):
            return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                    and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

        def __repr__(self):
            return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

    class Env:
        metadata = {"render.modes": []}
        reward_range = (-float("inf"), float("inf"))
        action_space = None
        observation_space = None

        def step(self, action):
            raise NotImplementedError

        def reset(self):
            raise NotImplementedError

        def render(self, mode="human"):
            raise NotImplementedError

        def close(self):
            pass

        def seed(self, seed=None):
            return None

        @property
        def unwrapped(self):
            return self

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            self.close()
            return False

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env
            self.action_space = env.action_space
            self.observation_space = env.observation_space
            self.reward_range = getattr(env, "reward_range", Env.reward_range)
            self.metadata = getattr(env, "metadata", Env.metadata)

        def __getattr__(self, name):
            if name.startswith("_") or name == "env":
                raise AttributeError(name)
            return getattr(self.env, name)

        def step(self, action):
            return self.env.step(action)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        def render(self, mode="human", **kwargs):
            return self.env.render(mode, **kwargs)

        def close(self):
            return self.env.close()

        def seed(self, seed=None):
            return self.env.seed(seed)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def __repr__(self):
            return "<%s%r>" % (type(self).__name__, self.env)

    discrete.Discrete = Discrete
    box.Box = Box
    spaces.Space, spaces.Discrete, spaces.Box = Space, Discrete, Box
    spaces.discrete, spaces.box = discrete, box
    gym.Env, gym.Wrapper, gym.Space, gym.spaces = Env, Wrapper, Space, spaces
    gym.__version__ = "0.0-rogue_gym_b200-shim"
    gym.__rogue_gym_shim__ = True
    return {"gym": gym, "gym.spaces": spaces, "gym.spaces.discrete": discrete, "gym.spaces.box": box}


def _resolve():
    try:
        import gym
        from gym import spaces
        return gym, spaces
    except ImportError:
        pass
    try:
        import gymnasium as gym
        from gymnasium import spaces
        return gym, spaces
    except ImportError:
        pass
    mods = _build_shim()
    sys.modules.update(mods)  # so that `import gym` in user code and in the reference's tests resolves
    return mods["gym"], mods["gym.spaces"]


gym, spaces = _resolve()
Env, Wrapper = gym.Env, gym.Wrapper
IS_SHIM = bool(getattr(gym, "__rogue_gym_shim__", False))
