#!/usr/bin/env python
"""Device time per step over a long rollout, in buckets of 50 steps, with the event counters per bucket.
Shows how the workload drifts from the synchronised start to its stationary mix. Experiment script."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
from rogue_gym_python.rollout import Shard, synthetic_actions
n, steps, bucket = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 3000, 50
sh = Shard("{}", 0, n)
stream = torch.cuda.ExternalStream(sh.stream())
acts = np.stack([synthetic_actions(t, sh.env_ids) for t in range(steps)])
d = torch.from_numpy(acts).cuda()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps // bucket + 1)]
stats = []
ev[0].record(stream)
for t in range(steps):
    sh.step_device(d.data_ptr() + t * n)
    if (t + 1) % bucket == 0:
        ev[(t + 1) // bucket].record(stream)
        if (t + 1) % 500 == 0:
            stats.append(((t + 1), sh.stats()))
sh.sync()
ms = [ev[i].elapsed_time(ev[i + 1]) / bucket * 1e3 for i in range(steps // bucket)]
print("us/step per %d-step bucket:" % bucket, " ".join("%.0f" % m for m in ms))
prev = None
for t, st in stats:
    if prev:
        print(t, {k: (st[k] - prev[k]) // 500 for k in st})
    prev = st
