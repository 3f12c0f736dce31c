#!/usr/bin/env python
"""Splits the cost of a host-mirror pass: a pass right after a step (compare + PCIe stores) against a second pass with
nothing left to send (compare only), and an empty synced call (launch + sync latency). Experiment script."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
from rogue_gym_python.rollout import Shard, synthetic_actions
n, K = 65536, 400
sh = Shard("{}", 0, n)
acts = np.stack([synthetic_actions(t, sh.env_ids) for t in range(K)])
dacts = torch.from_numpy(acts).cuda()
L, h = sh.L, sh.h
obs, hist = sh.mirror(with_history=len(sys.argv) > 1)
nb = C.c_uint64()
for t in range(200):
    L.rg_step(h, dacts.data_ptr() + t * n, 1)
sh.quiesce(); sh.sync()
t_step = t_first = t_second = t_sync = 0.0
sent = 0
for t in range(200, K):
    a = time.perf_counter(); L.rg_step(h, dacts.data_ptr() + t * n, 1); L.rg_sync(h)
    b = time.perf_counter(); L.rg_mirror_sync(h, C.byref(nb)); sent += nb.value
    c = time.perf_counter(); L.rg_mirror_sync(h, C.byref(nb))
    d = time.perf_counter(); L.rg_sync(h)
    e = time.perf_counter()
    t_step += b - a; t_first += c - b; t_second += d - c; t_sync += e - d
m = K - 200
print("step+sync %.1f us | mirror pass with changes %.1f us (%.0f B) | mirror pass, nothing to send %.1f us | empty sync %.1f us"
      % (t_step / m * 1e6, t_first / m * 1e6, sent / m, t_second / m * 1e6, t_sync / m * 1e6))
