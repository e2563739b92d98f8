#!/usr/bin/env python3
"""Engine 17 of the Q-net conv kernel: clock64 stamps of CTA 0 (snk_qnet_debug_timing), 64 slots per iteration.

slots: 0 iteration start | 3 conv2 -> conv3 barrier passed | 6 conv1 of the next iteration done (warp 13)
       8+t issuer starts tile t | 16+t issuer has queued tile t | 24+t warp 6 sees tile t complete | 32+t warp 6 starts its columns
       40+t warp 6 releases the accumulator | 48+t warp 6 has issued the TMA store | 56+t warp 4 has parked its half
"""
import ctypes as C
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
buf = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
L = S.lib()
L.snk_qnet_debug_timing.argtypes = [C.c_void_p, C.c_void_p]
for _ in range(2):
    net(obs)
L.snk_qnet_debug_timing(net._q, C.c_void_p(buf.data_ptr()))
net(obs)
torch.cuda.synchronize()
t = buf.cpu().view(8, 64).tolist()
for it in range(2, 5):
    r = t[it]
    z = r[0]
    rel = lambda k: (r[k] - z) if r[k] else None
    print("iter %d: total %d | barrier at %s | conv1 done at %s" % (it, t[it + 1][0] - z, rel(3), rel(6)))
    for name, base in [("issuer starts tile", 8), ("issuer queued tile", 16), ("w6 sees tile full", 24), ("w6 starts columns", 32),
                       ("w6 releases acc", 40), ("w6 issued store", 48), ("w4 parked", 56)]:
        print("   %-20s" % name, " ".join("%6s" % rel(base + k) for k in range(6)))
