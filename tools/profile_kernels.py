#!/usr/bin/env python3
"""Launches every kernel of the library a few times at its benchmark size, for ONE ncu capture:

  python tools/profile_kernels.py                                    # must exit 0 on its own first
  ncu --set full --clock-control none --import-source on -k regex:'k_step|k_rollout|k_qnet|k_gram|k_sample_grads|k_center' \
      -o gpurun_out/r02_kernels python tools/profile_kernels.py

Order of launches (the summaries under profiles/ refer to it): k_step F32+select x2, I8 x2, PACKED2 x2, BITS x2, NONE x2 at 2^20 envs;
k_rollout_ws (4,096 envs x 200 steps) x2; Q-net forward f32 x2 and bf16 x2 at 65,536 samples; k_sample_grads (1,024 samples)
x2; centre + pack(f64) + pack(f32) + Gram terms 3 and 1 at K=1000 x2; one 6,250 x 6,250 block Gram (the config-5b shard block)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

S = g.load_package()
dev = torch.device("cuda", 0)
REPS = 1 if os.environ.get("PROFILE_ONCE") else 2          # under ncu one launch of each is enough (it replays the kernel itself)
E = 1 << 20
env = S.SnakeGame(E, auto_reset=True)
q = torch.rand(E, 3, device=dev) * 2 - 1
u = torch.rand(E, device=dev)
r = torch.randint(0, 3, (E,), device=dev, dtype=torch.uint8)
for fmt in ("f32", "i8", "packed2", "bits", None):
    out = env.alloc_outputs(obs=fmt, mask=True, act=True)
    for _ in range(REPS):
        env.step_fused(q=q, eps=0.05, u=u, ridx=r, out=out)
    del out
torch.cuda.synchronize()
env.close()

env2 = S.SnakeGame(4096, auto_reset=True)
acts = torch.randint(0, 3, (200, 4096), device=dev, dtype=torch.uint8)
ro = env2.rollout(acts, obs="f32", mask=True)
if REPS > 1:
    env2.rollout(acts, out=ro)
torch.cuda.synchronize()

n = 65536
env3 = S.SnakeGame(n, auto_reset=True)
o3 = env3.alloc_outputs(obs="f32", mask=True)
for t in range(10):
    env3.step_fused(act_idx=torch.randint(0, 3, (n,), device=dev, dtype=torch.uint8), out=o3)
layers = S.qnet.glorot_layers(0)
nets = {p: S.qnet.QNet(layers, dev, p) for p in ("f32", "bf16")}
for p in ("f32", "bf16"):
    for _ in range(REPS):
        nets[p](o3["obs"])
torch.cuda.synchronize()

B = 1024
st = o3["obs"][:B].contiguous()
ac = torch.randint(0, 3, (B,), device=dev, dtype=torch.uint8)
y = torch.randn(B, device=dev, dtype=torch.float64)
plan_g = S.GramPlan(B, S.qnet.N_PARAMS, dev)
for _ in range(REPS):
    nets["f32"].sample_grads(st, ac, y, planes=plan_g.planes(), want_loss=False)
torch.cuda.synchronize()

K, P = 1000, 181395
A = torch.randn(K, P, device=dev, dtype=torch.float32).double()
for _ in range(REPS):
    S.center_columns(A)
plan = S.GramPlan(K, P, dev)
A32 = A.float()
G = torch.empty(K, K, dtype=torch.float32, device=dev)
for _ in range(REPS):
    plan.pack(A)
    plan.pack(A32)
    plan.gram(3, 0, out=G)
    plan.gram(1, 0, out=G)
torch.cuda.synchronize()
del A, A32, plan

R = 6250
J = torch.randn(R, P, device=dev, dtype=torch.float32)
plan5 = S.GramPlan(R, P, dev).pack(J)
G5 = torch.empty(R, R, dtype=torch.float32, device=dev)
plan5.gram(3, 0, out=G5)
torch.cuda.synchronize()
print("ok")
