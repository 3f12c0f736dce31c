#!/usr/bin/env python
"""Lone-warp latency of one floor generation per seed (k_reset on a batch of one env), with the number of mazes."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from rogue_gym_python import _cabi, _rogue_gym
from helpers import gpu_dump
level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = _rogue_gym.GameState(1000, "{}")
b = g._batch
stream = torch.cuda.ExternalStream(b.L.rg_stream(b.h))
rows = []
for seed in range(1, 241):
    g.set_seed(seed)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.L.rg_sync(b.h)
    e0.record(stream)
    b.L.rg_reset(b.h)
    e1.record(stream)
    b.L.rg_sync(b.h)
    d = gpu_dump(b, 0, with_maps=False)
    kinds = d["rooms"][:, 0]
    rows.append((e0.elapsed_time(e1) * 1e3, int((kinds == 1).sum()), int((kinds == 2).sum()), seed))
a = np.array(rows)
print("us: min %.0f median %.0f p90 %.0f max %.0f" % (a[:, 0].min(), np.median(a[:, 0]), np.percentile(a[:, 0], 90), a[:, 0].max()))
for m in range(0, 5):
    sel = a[a[:, 1] == m]
    if len(sel):
        print("mazes=%d: n=%3d median %.0f us max %.0f" % (m, len(sel), np.median(sel[:, 0]), sel[:, 0].max()))
print("slowest:", sorted(rows)[-5:])
