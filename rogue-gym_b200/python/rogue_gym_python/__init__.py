"""Drop-in for the reference's `rogue_gym_python` package (python/rogue_gym_python in the wheel
built by the reference's python/setup.py:57): the extension module `_rogue_gym` is provided by
the B200 C-ABI library instead of the PyO3/Rust build."""
from . import _rogue_gym  # noqa: F401
