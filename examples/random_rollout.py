#!/usr/bin/env python
"""Three ways to drive the simulator, from the reference-shaped API to the device-resident one.

    python examples/random_rollout.py            (needs a B200; build first: python __graft_entry__.py)

1. `RogueEnv`          one game, gym-style, PlayerState objects       (reference: rogue_gym.envs.RogueEnv)
2. `ParallelRogueEnv`  N games in lockstep with auto-reset, lists     (reference: rogue_gym.envs.ParallelRogueEnv)
3. `DeviceRogueEnv`    N games, CUDA tensors in and out, no host trip (new: for trainers)
plus `ParallelGameState.step_arrays`: N games, numpy views of a host mirror that the device keeps current.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from rogue_gym.envs import DeviceRogueEnv, DungeonType, ImageSetting, ParallelRogueEnv, RogueEnv, StatusFlag  # noqa: E402

# 1 ---------------------------------------------------------------- one game
env = RogueEnv(seed=1)
state, reward, done, _ = env.step("hjkl")          # a string of keys, or an index into RogueEnv.ACTIONS
print(env)                                          # the 80x24 screen and the status line
print("gold", state.gold, "level", state.dungeon_level, "image", env.state_to_image(state).shape)

# 2 ---------------------------------------------------------------- N games, list API
penv = ParallelRogueEnv([{"seed": s} for s in range(1, 9)], max_steps=100)
states, rewards, dones, _ = penv.step([1, 2, 3, 4, 5, 6, 7, 8])
print("list API:", len(states), "states, rewards", rewards)
penv.close()

# 2b --------------------------------------------------------------- N games, numpy views of the host mirror
game = ParallelRogueEnv([{}] * 4096).game               # rogue_gym_python._rogue_gym.ParallelGameState
game.seed(list(range(1, 4097)))
game.reset()
keys = np.frombuffer(b".hjklnbuy>s", np.uint8)
t0 = time.time()
for _ in range(200):
    obs = game.step_arrays(keys[np.random.randint(0, 11, size=4096)])   # dict of numpy views, refreshed in place
print("host mirror: screen", obs["screen"].shape, "done", int(obs["done"].sum()),
      "| %.2f M env-steps/s through numpy" % (4096 * 200 / (time.time() - t0) / 1e6))
game.close()

# 3 ---------------------------------------------------------------- N games on the device
n = 16384
denv = DeviceRogueEnv({}, num_envs=n, image_setting=ImageSetting(DungeonType.GRAY, StatusFlag.HP_CURRENT, True),
                      seeds=range(1, n + 1), stair_reward=50.0)
obs = denv.reset()                                       # CompactObs on cuda:0 (symbol ids u8 [N,24,80], status i32 [N,9], visited map), rewritten in place by step()
torch.cuda.synchronize()
t0 = time.time()
total = torch.zeros((), device="cuda")
for _ in range(300):
    actions = torch.randint(0, denv.ACTION_LEN, (n,), device="cuda")   # a policy would look at `obs` here
    obs, reward, done, _ = denv.step(actions)
    total += reward.sum()
torch.cuda.synchronize()
image = denv.expand(obs)                                 # float32 [N, 3, 24, 80]: the reference's ImageSetting.expand, on the device
print("device env: obs", tuple(obs.symbols.shape), "-> image", tuple(image.shape),
      "| %.1f M env-steps/s incl. observation | reward sum %.0f | errors %d"
      % (n * 300 / (time.time() - t0) / 1e6, float(total), int((denv.errors() != 0).sum())))
denv.close()
