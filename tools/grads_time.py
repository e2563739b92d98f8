#!/usr/bin/env python3
"""snk_qnet_sample_grads at the config-5b shard size (6,250 transitions of a real rollout), device-timed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

S = g.load_package()
dev = torch.device("cuda", 0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 6250
env = S.SnakeGame(16384, auto_reset=True)
ring = S.ReplayBuffer(capacity=50000)
net = S.qnet.QNet(S.qnet.glorot_layers(0), dev, "f32")
ro = S.rollout.Rollout(env, net, net, ring, epsilon=0.3)
for _ in range(4):
    ro.step()
batch = ring.stack_exp(ring.sample_indices(R))
y = S.masked_target(net(batch["next_states"]), batch["mask"], batch["rewards"], batch["dones"])
plan = S.GramPlan(R, S.qnet.N_PARAMS, dev)
for _ in range(2):
    net.sample_grads(batch["states"], batch["actions"], y, planes=plan.planes(), want_loss=False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    net.sample_grads(batch["states"], batch["actions"], y, planes=plan.planes(), want_loss=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("sample_grads: %d samples in %.2f ms = %.2f us per sample; %.1f GFLOP/s useful FP32 (14.6 MFLOP per sample); plane write %.0f GB/s"
      % (R, ms, 1e3 * ms / R, 14.6e6 * R / (ms * 1e-3) / 1e9, R * 181440 * 4 / (ms * 1e-3) / 1e9))
