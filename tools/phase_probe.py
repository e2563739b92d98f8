#!/usr/bin/env python3
"""Cycle-level phase timings of two latency-bound kernels (CTA 0): the Float32-faithful Q-net conv kernel (clock64 stamps per
iteration) and the warp-specialised small-batch rollout kernel (cycle sums of the logic warp, the mask warp and expansion thread 0)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

S = g.load_package()
L = S.lib()
n = 65536
env = S.SnakeGame(n)
obs = env.assemble_state("f32")
net = S.qnet.QNet(S.qnet.glorot_layers(0), env.device, precision="f32")
for _ in range(2):
    net(obs)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
L.snk_qnet_debug_timing(net._q, C.c_void_p(buf.data_ptr()))
net(obs)
torch.cuda.synchronize()
L.snk_qnet_debug_timing(net._q, None)
t = buf.cpu().view(-1, 8)[:12]
print("split engine, CTA 0, cycles relative to the iteration's start stamp")
print("it   conv2_done  after_barrier  c3_full(MMAs done)  epilogue_done  iter_end   conv1_done   next_it_start-this")
for i in range(1, 10):
    r = t[i]
    print(i, [int(r[k] - r[0]) for k in (1, 2, 3, 4, 5, 6)], int(t[i + 1][0] - r[0]))

env2 = S.SnakeGame(4096)
T = 200
acts = torch.randint(0, 3, (T, 4096), device="cuda", dtype=torch.uint8)
out = env2.rollout(acts, obs="f32", mask=True)
pb = torch.zeros(8, dtype=torch.int64, device="cuda")
L.snk_debug_rollout_timing.argtypes = [C.c_void_p]
L.snk_debug_rollout_timing(C.c_void_p(pb.data_ptr()))
env2.rollout(acts, out=out)
torch.cuda.synchronize()
L.snk_debug_rollout_timing(None)
p = pb.cpu().tolist()
print("rollout_ws, CTA 0, cycles per step: logic warp total %.0f (waiting for the consumers %.0f, its own chain %.0f) | mask warp: waiting %.0f, "
      "mask + scalars %.0f | expansion warps: waiting %.0f, boards + expansion %.0f   (per observation format: tools/rollout_probe.py)"
      % tuple(p[k] / T for k in (0, 1, 6, 2, 3, 5, 4)))
