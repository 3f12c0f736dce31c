#!/usr/bin/env python
"""exp_host_rate.py on every GPU at once (torchrun): per rank, wall time of the enqueue loop alone against the device
time of the same steps - is the device-resident rollout host-bound when 8 processes share the host? Experiment."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
import torch.distributed as dist
from rogue_gym_python.rollout import Shard, synthetic_actions
rank, world = int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if os.environ.get("PIN", "1") == "1" and world > 1:
    cores = sorted(os.sched_getaffinity(0)); per = max(1, len(cores) // world)
    os.sched_setaffinity(0, cores[rank * per:(rank + 1) * per])
n, warm, steps = 65536, 900, 600
sh = Shard("{}", rank * n, (rank + 1) * n, device=rank)
stream = torch.cuda.ExternalStream(sh.stream())
acts = np.stack([synthetic_actions(t, sh.env_ids) for t in range(steps + warm)])
d = torch.from_numpy(acts).cuda()
torch.cuda.synchronize()
for t in range(warm):
    sh.step_device(d.data_ptr() + t * n)
sh.quiesce(); sh.sync()
if world > 1:
    dist.init_process_group("gloo"); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record(stream)
for t in range(warm, warm + steps):
    sh.step_device(d.data_ptr() + t * n)
t1 = time.perf_counter()
sh.quiesce()
e1.record(stream)
sh.sync()
print("rank %d: host enqueue %.1f us/step, device %.1f us/step" % (rank, (t1 - t0) / steps * 1e6, e0.elapsed_time(e1) / steps * 1e3), flush=True)
if world > 1:
    dist.barrier()
