#!/usr/bin/env python3
"""Minimal native Q-net forward run for ncu (65,536 samples, 3 forwards)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

S = graft.load_package()
n = 65536
env = S.SnakeGame(n, auto_reset=True)
obs = env.assemble_state("f32")
net = S.qnet.QNet(S.qnet.glorot_layers(0), env.device, precision="bf16")
for _ in range(3):
    q = net(obs)
torch.cuda.synchronize()
print("ok", q[0].tolist())
