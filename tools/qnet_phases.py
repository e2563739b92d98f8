#!/usr/bin/env python3
"""Per-phase cycle counts of the Q-net conv kernel of engines 12 / 16 (CTA 0) via snk_qnet_debug_timing.
SNK_QNET_ENGINE=12 (default here) or 16 with RAW=1; engine 17, the library default, has its own tool: qnet_phases17.py."""
import ctypes as C
import os
import sys

import torch

os.environ.setdefault("SNK_QNET_ENGINE", "12")

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

S = graft.load_package()
n = 65536
env = S.SnakeGame(n, auto_reset=True)
obs = env.assemble_state("f32")
net = S.qnet.QNet(S.qnet.glorot_layers(0), env.device, backend="native")
buf = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
L = S.lib()
L.snk_qnet_debug_timing.argtypes = [C.c_void_p, C.c_void_p]
for _ in range(2):
    net(obs)
L.snk_qnet_debug_timing(net._q, C.c_void_p(buf.data_ptr()))
net(obs)
torch.cuda.synchronize()
t = buf.cpu().view(-1, 8)
names = ["obs->A0", "conv1", "conv2", "conv3 mma (wait c3_full)", "conv3 epilogue+sync"]
for it in range(1, 6):
    row = t[it]
    print("iter %d:" % it, ", ".join("%s %d" % (names[k], int(row[k + 1] - row[k])) for k in range(5)),
          "| next conv1 done at +%d of the conv3 phase" % int(row[6] - row[3]), "| seq conv1 %d" % int(row[7] - row[5]), "| total", int(t[it + 1][0] - row[0]))

if os.environ.get("RAW"):
    # engine 16 stamps: 0 start, 3 conv2 done, 4 c3_full seen, 1 upper halves dumped, 2 warp 4 finalised, 5 end barrier, 6 conv1 done, 7 end
    for it in range(1, 5):
        r = [int(x) for x in t[it]]
        print("raw iter %d: conv2 %d | conv3 mma %d | dump+bar %d | finalise(warp4) %d | to end barrier %d | conv1 done +%d | tail %d | total %d" % (
            it, r[3] - r[0], r[4] - r[3], r[1] - r[4], r[2] - r[1], r[5] - r[2], r[6] - r[3], r[7] - r[1], int(t[it + 1][0]) - r[0]))

if os.environ.get("RAW2"):
    for it in range(1, 4):
        r = [int(x) for x in t[it]]
        print("raw2 iter %d: c3_full->bar %d | bar->first tmem load done %d | scratch loads done %d | rest to end barrier %d" % (it, r[1] - r[4], r[7] - r[1], r[2] - r[7], r[5] - r[2]))
