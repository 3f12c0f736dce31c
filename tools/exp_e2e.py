#!/usr/bin/env python
"""Where the host-facing step spends its time: per-step wall clock of the C-ABI calls, synced every step."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
from rogue_gym_python import _cabi
from rogue_gym_python.rollout import Shard, synthetic_actions
n, K = int(os.environ.get("E2E_N", "65536")), 600
sh = Shard("{}", 0, n)
acts = np.stack([synthetic_actions(t, sh.env_ids) for t in range(K)])
hacts = torch.from_numpy(acts).pin_memory()
dacts = hacts.cuda()
L, h = sh.L, sh.h
def run(name, fn):
    sh.reseed_and_reset()
    for t in range(200): fn(t)
    sh.sync()
    t0 = time.perf_counter()
    for t in range(200, K): fn(t)
    sh.sync()
    print("%-34s %7.1f us/step" % (name, (time.perf_counter() - t0) / (K - 200) * 1e6), flush=True)
run("rg_step (device actions), no sync", lambda t: L.rg_step(h, dacts.data_ptr() + t * n, 1))
def f(t):
    L.rg_step(h, dacts.data_ptr() + t * n, 1); L.rg_sync(h)
run("rg_step + rg_sync", f)
obs, hist = sh.mirror()
nb = C.c_uint64()
run("rg_step_mirror", lambda t: L.rg_step_mirror(h, hacts.data_ptr() + t * n, 1, C.byref(nb)))
def g(t):
    L.rg_step(h, dacts.data_ptr() + t * n, 1); L.rg_mirror_sync(h, C.byref(nb))
run("rg_step(dev) + rg_mirror_sync", g)
