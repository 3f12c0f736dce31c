#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel of the path on a few hundred envs (reset, step with
auto-reset and descents, MoveUntil keys, prefetch, host mirror, encoders, dump)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from rogue_gym_python import _rogue_gym
from helpers import KEYS19, gpu_dump
n, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 192, int(sys.argv[2]) if len(sys.argv) > 2 else 60
cfgs = [{}, {"width": 50, "height": 19, "dungeon": {"style": "rogue", "room_num_x": 2, "room_num_y": 2}}]
rng = np.random.RandomState(3)
for cfg in cfgs:
    pg = _rogue_gym.ParallelGameState(20, [json.dumps(cfg)] * n)
    pg.seed(list(range(1, n + 1)))
    pg.reset()
    b = pg._batch
    for t in range(steps):
        keys = KEYS19[rng.randint(0, len(KEYS19), size=n)]
        if t % 2:
            b.step_mirror(keys, True)
        else:
            try:
                b.step(keys, True)
            except RuntimeError:
                pass
    st = pg.states()[0]
    st.symbol_image(0x1FF); st.gray_image_with_hist(3)
    gpu_dump(b, 0, with_maps=True)
    pg.close()
print("sanitize workload done")
