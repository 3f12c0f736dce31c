#!/usr/bin/env python
"""Is the device-resident rollout host-bound? Wall time of the enqueue loop alone (no sync) against the device time
of the same steps. Experiment script."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
from rogue_gym_python.rollout import Shard, synthetic_actions
n, steps, warm = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 1500, 500
sh = Shard("{}", 0, n)
stream = torch.cuda.ExternalStream(sh.stream())
acts = np.stack([synthetic_actions(t, sh.env_ids) for t in range(steps + warm)])
d = torch.from_numpy(acts).cuda()
torch.cuda.synchronize()
for t in range(warm):
    sh.step_device(d.data_ptr() + t * n)
sh.quiesce(); sh.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record(stream)
for t in range(warm, warm + steps):
    sh.step_device(d.data_ptr() + t * n)
t1 = time.perf_counter()
sh.quiesce()
e1.record(stream)
sh.sync()
t2 = time.perf_counter()
print("host enqueue %.1f us/step, wall %.1f us/step, device %.1f us/step" % ((t1 - t0) / steps * 1e6, (t2 - t0) / steps * 1e6, e0.elapsed_time(e1) / steps * 1e3))
print(sh.stats())
