#!/usr/bin/env python3
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
S = g.load_package()
env = S.SnakeGame(65536); obs = env.assemble_state("f32")
net = S.qnet.QNet(S.qnet.glorot_layers(0), env.device, backend="native")
for _ in range(3): net(obs)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): net(obs)
e1.record(); torch.cuda.synchronize()
print("qnet forward ms", e0.elapsed_time(e1) / 20)
