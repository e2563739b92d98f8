#!/usr/bin/env python3
"""Minimal Gram run for ncu: pack once, run snk_gram (terms 3 then 1) a few times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

S = graft.load_package()
dev = torch.device("cuda", 0)
K, P = 1000, 181395
A = torch.randn(K, P, device=dev, dtype=torch.float32)
plan = S.GramPlan(K, P, dev).pack(A)
G = torch.empty(K, K, dtype=torch.float32, device=dev)
for terms in (3, 1):
    for _ in range(3):
        plan.gram(terms, 0, out=G)
torch.cuda.synchronize()
print("ok", float(G[0, 0]))
