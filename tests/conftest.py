import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CONFIGS, DEFAULT, MINI, MINI_NOMON  # noqa: E402,F401  (shared constants live in helpers.py)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fixtures():
    with open(os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gif_frames():
    with open(os.path.join(ROOT, "tests", "golden", "reference_gif_frames.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def cabi():
    """The product's C-ABI library; built in-tree if the .so is missing (nvcc cross-compiles on CPU)."""
    from rogue_gym_python import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "rogue-gym_b200"), "-s", "-j4"])
    _cabi.lib()
    return _cabi


@pytest.fixture(scope="session")
def gpu(cabi):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rogue_gym_python import _rogue_gym
    return _rogue_gym
