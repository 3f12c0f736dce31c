"""Drop-in for the reference's `rogue_gym` Python package (python/rogue_gym/__init__.py): the gym
layer (`rogue_gym.envs`) on top of the B200 extension-module mirror `rogue_gym_python._rogue_gym`."""
from . import envs  # noqa: F401

try:  # optional trainer adapter, only when `rainy` is installed (python/rogue_gym/__init__.py:3-7)
    import rainy  # noqa: F401

    from . import rainy_impls  # noqa: F401
except ImportError:
    pass

__version__ = "0.0.2+b200"
