#!/usr/bin/env python
"""Host-facing step (rg_step_mirror, synced every step) on every GPU of the box at once: one process per GPU under
torchrun, each prints its own us/step. Shows what the shared host costs. Experiment script.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/exp_e2e_multi.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
from rogue_gym_python.rollout import Shard, synthetic_actions
rank, world = int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if os.environ.get("PIN", "1") == "1" and world > 1:
    cores = sorted(os.sched_getaffinity(0)); per = max(1, len(cores) // world)
    os.sched_setaffinity(0, cores[rank * per:(rank + 1) * per])
n, burn, K = 65536, 600, 300
sh = Shard("{}", rank * n, (rank + 1) * n, device=rank)
acts = np.stack([synthetic_actions(t, sh.env_ids) for t in range(burn + K)])
hacts = torch.from_numpy(acts).pin_memory()
dacts = hacts.cuda()
for t in range(burn):
    sh.step_device(dacts.data_ptr() + t * n)
sh.quiesce(); sh.sync()
sh.mirror()
import torch.distributed as dist
if world > 1:
    dist.init_process_group("gloo")
    dist.barrier()
t0 = time.perf_counter()
sent = 0
for t in range(burn, burn + K):
    sent += sh.step_mirror(hacts.data_ptr() + t * n)
dt = (time.perf_counter() - t0) / K * 1e6
print("rank %d: %.1f us/step, %.0f B/step to host" % (rank, dt, sent / K), flush=True)
if world > 1:
    dist.barrier()
