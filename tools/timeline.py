#!/usr/bin/env python
"""Prints the in-kernel timeline (RG_TRACE=1) of a few steady-state steps of the bench workload."""
import ctypes as C, json, os, sys
os.environ["RG_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
from rogue_gym_python import _cabi
from rogue_gym_python.rollout import Shard, synthetic_actions
n, steps = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 300
sh = Shard("{}", 0, n)
acts = np.stack([synthetic_actions(t, sh.env_ids) for t in range(steps)])
d = torch.from_numpy(acts).cuda()
out = np.zeros((512, 12, 2), np.uint64)
launched = C.c_int64()
MIRROR = os.environ.get("TL_MIRROR") == "1"  # time the host-facing step (rg_step_mirror, synced every step) instead
if MIRROR:
    hacts = torch.from_numpy(acts).pin_memory()
    sh.mirror()
for t in range(steps):
    if t == steps - 200:  # the slots wrap every 512 steps: clear them shortly before the end
        sh.sync()
        _cabi.check(sh.L.rg_trace(sh.h, out.ctypes.data, C.byref(launched)), sh.h)
    if MIRROR:
        sh.step_mirror(hacts.data_ptr() + t * n)
    else:
        sh.step_device(d.data_ptr() + t * n)
sh.sync()
_cabi.check(sh.L.rg_trace(sh.h, out.ctypes.data, C.byref(launched)), sh.h)
names = ["playerA", "monstersA", "fast", "full", "resets", "prefetch", "playerB", "monstersB", "mirror1", "mirror2", "scan"]
last = (launched.value - 1) % 512
order = [(last - k) % 512 for k in range(12, 0, -1)]
t0 = int(out[order[0], 0, 0])
for s in order:
    row = []
    for k, nm in enumerate(names):
        a, b = int(out[s, k, 0]), int(out[s, k, 1])
        if b == 0 or a == 2**64 - 1:
            continue
        row.append("%s %7.1f..%7.1f (%5.1f)" % (nm, (a - t0) / 1e3, (b - t0) / 1e3, (b - a) / 1e3))
    print("slot %2d: " % s + " | ".join(row))
print(sh.stats())
