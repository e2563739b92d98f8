#!/usr/bin/env python3
"""Small run of every kernel of the library, for compute-sanitizer --tool memcheck (ragged sizes on purpose):
  compute-sanitizer --tool memcheck python tools/sanitize_env.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

S = graft.load_package()
n = 1000 + 37
env = S.SnakeGame(n, auto_reset=True, food_list=[(3, 5), (5, 5), (2, 2), (9, 9)])
rb = S.ReplayBuffer(capacity=777)
rng = np.random.default_rng(0)
for fmt in ("f32", "i8", "i64", "packed2", "bits"):
    out = env.alloc_outputs(obs=fmt, mask=True, ep_stats=True, act=True)
    for t in range(6):
        q = torch.rand(n, 3, device="cuda")
        env.step_fused(q=q, eps=0.3, out=out, replay=rb)
        env.step_fused(act_idx=torch.from_numpy(rng.integers(0, 3, n).astype(np.uint8)).cuda(), out=out)
    st = env.assemble_state(fmt)
    env.patch_reset_obs(out["done"], st, fmt)
env.step(torch.zeros(n, dtype=torch.uint8, device="cuda"))
env.step_abs(torch.full((n,), 3, dtype=torch.uint8, device="cuda"))
env.virtual_step(); env.available_actions(); env.epsilon_greedy(torch.rand(n, 3, device="cuda"), 0.5)
env.score; env.lost; env.error_flags; env.steps; env.count_errors()
acts = torch.from_numpy(rng.integers(0, 3, (9, n)).astype(np.uint8)).cuda()
env.rollout(acts, obs="f32", mask=True, ep_stats=True)
small = S.SnakeGame(45, auto_reset=False)
small.rollout(acts[:, :45].contiguous(), obs="packed2", mask=True)
b = rb.stack_exp(rb.sample_indices(333), ep_stats=True)
S.masked_target(torch.rand(n, 3, device="cuda"), out["mask"], out["reward"], out["done"])
D = torch.randn(17, 1031, dtype=torch.float64, device="cuda")
S.center_columns(D)
host = {"obs_fmt": "f32", "act_idx": S.pinned_empty((n,), torch.uint8), "reward": S.pinned_empty((n,), torch.float32),
        "done": S.pinned_empty((n,), torch.uint8), "obs": S.pinned_empty((n, 2, 10, 10), torch.float32),
        "mask": S.pinned_empty((n, 3), torch.uint8)}
host["act_idx"].zero_()
env.step_fused_host(host)
env.sync()
hb = {"obs_fmt": "bits", "q": S.pinned_empty((n, 3), torch.float32), "obs": S.pinned_empty((n, 24), torch.uint8)}
hb["q"].copy_(torch.rand(n, 3))
env.step_fused_host(hb, q=True, eps=0.1, replay=rb)
env.sync()
S.unpack_bits(hb["obs"])
# tensor-core kernels and the gradient kernel at small, ragged sizes
A = torch.randn(300, 1031, device="cuda", dtype=torch.float64)
for cg in (1, 2):
    S.lib().snk_gram_config(cg)
    plan = S.GramPlan(300, 1031, A.device).pack(A)
    plan.gram(3); plan.gram(1); plan.gram(3, 32)
S.lib().snk_gram_config(2)
from snake_b200 import gram_sharded as GS
peers = GS.LocalPeers(301, 515, 3, torch.device("cuda", 0))
peers.gram(A[:, :515].float()[:301].contiguous() if A.shape[0] >= 301 else torch.randn(301, 515, device="cuda"))
peers.free()
peers = GS.LocalPeers(200, 515, 2, torch.device("cuda", 0))
peers.gram(torch.randn(200, 515, device="cuda"))
peers.free()
layers = S.qnet.glorot_layers(0)
for prec in ("f32", "bf16"):
    net = S.qnet.QNet(layers, "cuda:0", precision=prec)
    net(torch.rand(37, 2, 10, 10, device="cuda"))
net = S.qnet.QNet(layers, "cuda:0", precision="f32")
B = 9
gp = S.GramPlan(B, S.qnet.N_PARAMS, "cuda:0")
net.sample_grads(torch.rand(B, 2, 10, 10, device="cuda"), torch.randint(0, 3, (B,), device="cuda", dtype=torch.uint8),
                 torch.randn(B, device="cuda", dtype=torch.float64), planes=gp.planes(), want_J=True, want_loss=True)
gp.gram(3)
torch.cuda.synchronize()
print("sanitize run ok", float(b["rewards"].sum()))
