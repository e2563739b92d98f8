#!/usr/bin/env python3
"""Times snk_center_columns at the config-5a size (L2 flushed between iterations) and prints a checksum of the results.
Use SNAKE_B200_LIB to point at a tuning variant of the library."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

S = graft.load_package()
dev = torch.device("cuda", 0)
K, P = 1000, 181395
g = torch.Generator(device=dev); g.manual_seed(5)
A0 = torch.randn(K, P, device=dev, dtype=torch.float32, generator=g).double()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
A = A0.clone()
ts = []
for i in range(8):
    A.copy_(A0)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m, v = S.center_columns(A); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = sorted(ts[2:])
chk = int(A.view(torch.int64).sum().item()) ^ int(m.view(torch.int64).sum().item()) ^ int(v.view(torch.int64).sum().item())
print("%s center_ms median %.4f min %.4f  hbm_frac %.3f  checksum %x" % (os.environ.get("SNAKE_B200_LIB", "default"), ts[len(ts) // 2], ts[0],
      3 * 8.0 * K * P / (ts[len(ts) // 2] * 1e-3) / 1e9 / 6455.6, chk & 0xFFFFFFFFFFFF))

z1 = torch.randn(P, dtype=torch.float64, device=dev)
z2 = torch.randn(K, dtype=torch.float64, device=dev)
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w = S.laplace.sample_model_weights(m, v, A, z1, z2); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = sorted(ts[2:])
print("sample_model_weights ms median %.4f  hbm_frac %.3f" % (ts[len(ts) // 2], 8.0 * K * P / (ts[len(ts) // 2] * 1e-3) / 1e9 / 6455.6))
