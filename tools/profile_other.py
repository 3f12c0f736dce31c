#!/usr/bin/env python
"""Launches each kernel off the step path once or twice for tools/profile_r2.sh: reset, prefetch / skeleton passes (a few
steps), both encoders, the compact encoder and the host mirror passes."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
from rogue_gym_python import _cabi
from rogue_gym_python.rollout import Shard, synthetic_actions
n = 65536
sh = Shard("{}", 0, n)
L, h = sh.L, sh.h
acts = np.stack([synthetic_actions(t, sh.env_ids) for t in range(40)])
hacts = torch.from_numpy(acts).pin_memory()
dacts = hacts.cuda()
for t in range(30):
    sh.step_device(dacts.data_ptr() + t * n)
sh.quiesce(); sh.sync()
cc = C.c_int()
for mode, flag, hist in ((0, 0, 0), (1, 0x1FF, 0)):
    ch = L.rg_encode_channels(h, mode, flag, hist)
    out = torch.empty((n, ch, sh.H, sh.W), dtype=torch.float32, device="cuda")
    _cabi.check(L.rg_encode(h, mode, flag, hist, out.data_ptr(), C.byref(cc)), h)
    sh.sync()
    del out
sym = torch.empty((n, sh.H, sh.W), dtype=torch.uint8, device="cuda")
stat = torch.empty((n, 9), dtype=torch.int32, device="cuda")
_cabi.check(L.rg_encode_compact(h, sym.data_ptr(), stat.data_ptr(), None), h)
sh.sync()
sh.mirror()
for t in range(30, 34):
    sh.step_mirror(hacts.data_ptr() + t * n)
print("done")
