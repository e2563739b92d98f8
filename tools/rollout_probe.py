#!/usr/bin/env python3
"""Config 2 (4,096 envs x 200 steps) through snk_rollout_fused per observation format, with the cycle counters of CTA 0
(snk_debug_rollout_timing): is the step time the logic warp's chain, the expanders' work, or their contention?"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

S = g.load_package()
n, T = 4096, 200
env = S.SnakeGame(n, auto_reset=True)
acts = torch.randint(0, 3, (T, n), device="cuda", dtype=torch.uint8)
prof = torch.zeros(8, dtype=torch.int64, device="cuda")
for fmt, mask in (("f32", True), ("i8", True), ("packed2", True), (None, True), (None, False)):
    out = env.rollout(acts, obs=fmt, mask=mask)
    for _ in range(3):
        env.rollout(acts, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        env.rollout(acts, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 / T * 1e3
    S.lib().snk_debug_rollout_timing(C.c_void_p(prof.data_ptr()))
    env.rollout(acts, out=out)
    torch.cuda.synchronize()
    S.lib().snk_debug_rollout_timing(None)
    p = [int(x) // T for x in prof.tolist()]
    print("%s obs %-8s mask %-5s: %.3f us per step = %.3g env-steps/s | cycles per step: logic %d (waiting %d, work %d, arrive %d) | mask warp: waiting %d, mask+scalars %d | expansion warps: waiting %d, boards+expansion %d"
          % (os.environ.get("SNAKE_B200_LIB", "default")[-16:], fmt, mask, us, n / us * 1e6, p[0], p[1], p[6], p[7], p[2], p[3], p[5], p[4]))
