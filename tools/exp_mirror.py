import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
import numpy as np, torch
from rogue_gym_python.rollout import Shard, synthetic_actions
n, K = 65536, 260
sh = Shard("{}", 0, n)
acts = torch.from_numpy(np.stack([synthetic_actions(t, sh.env_ids) for t in range(K)])).pin_memory()
obs, hist = sh.mirror()
for t in range(K):
    sh.step_mirror(acts.data_ptr() + t * n)
print("done")
