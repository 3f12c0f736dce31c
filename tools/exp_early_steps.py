#!/usr/bin/env python
"""Per-step device time of the first steps after a reset (CUDA events around every rg_step), to see what a short
bench window (--steps 20 --warmup 5) measures. Experiment script."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
from rogue_gym_python.rollout import Shard, synthetic_actions
n, steps = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 80
sh = Shard("{}", 0, n)
stream = torch.cuda.ExternalStream(sh.stream())
acts = np.stack([synthetic_actions(t, sh.env_ids) for t in range(steps)])
d = torch.from_numpy(acts).cuda()
torch.cuda.synchronize()
for rep in range(2):
    sh.reseed_and_reset()
    sh.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    prev = sh.stats()
    ev[0].record(stream)
    for t in range(steps):
        sh.step_device(d.data_ptr() + t * n)
        if t == 4:
            sh.quiesce()
        ev[t + 1].record(stream)
    sh.sync()
    ms = [ev[t].elapsed_time(ev[t + 1]) for t in range(steps)]
    print("rep", rep, " ".join("%.3f" % m for m in ms))
    st = sh.stats()
    print({k: st[k] - prev[k] for k in st})
