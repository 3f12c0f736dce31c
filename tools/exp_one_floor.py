import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))
from rogue_gym_python import _rogue_gym
g = _rogue_gym.GameState(1000, "{}")
g.set_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 139)
for _ in range(4):
    g.reset()
print("ok")
