#!/usr/bin/env python3
"""Minimal snk_center_columns + snk_gram_pack run for ncu (config 5a size: Float64 D, K=1000 snapshots x P=181,395)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

S = graft.load_package()
dev = torch.device("cuda", 0)
K, P = 1000, 181395
A = torch.randn(K, P, device=dev, dtype=torch.float32).double()
for _ in range(3):
    S.center_columns(A)
plan = S.GramPlan(K, P, dev)
for _ in range(2):
    plan.pack(A)
torch.cuda.synchronize()
print("ok", float(A[0, 0]))
