#!/usr/bin/env python3
"""Minimal multi-step rollout run for ncu: 4,096 envs x 200 steps, three launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

S = graft.load_package()
env = S.SnakeGame(4096, auto_reset=True)
acts = torch.randint(0, 3, (200, 4096), device="cuda", dtype=torch.uint8)
out = env.rollout(acts, obs="f32", mask=True)
for _ in range(2):
    env.rollout(acts, out=out)
torch.cuda.synchronize()
print("ok")
