#!/usr/bin/env python3
"""snk_rollout_fused at BASELINE config 2 (4,096 envs x 200 steps, f32 obs + mask), device-timed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

S = g.load_package()
n, T = 4096, 200
env = S.SnakeGame(n, auto_reset=True)
acts = torch.randint(0, 3, (T, n), device="cuda", dtype=torch.uint8)
out = env.rollout(acts, obs="f32", mask=True)
for _ in range(3):
    env.rollout(acts, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    env.rollout(acts, out=out)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 / T * 1e3
print("%s: %.3f us per step = %.3g env-steps/s" % (os.environ.get("SNAKE_B200_LIB", "default"), us, n / us * 1e6))
