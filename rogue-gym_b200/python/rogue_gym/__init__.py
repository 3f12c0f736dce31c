"""Drop-in for the reference's `rogue_gym` Python package (python/rogue_gym/__init__.py): the gym
layer (`rogue_gym.envs`) on top of the B200 extension-module mirror `rogue_gym_python._rogue_gym`."""
from . import envs  # noqa: F401

# (The reference also ships `rainy_impls.py`, an adapter for the third-party `rainy` trainer that expands one Python
# PlayerState at a time; it is off the accelerated path and not reproduced - trainers use rogue_gym.envs.DeviceRogueEnv.)

__version__ = "0.0.2+b200"
