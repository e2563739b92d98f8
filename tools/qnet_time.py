#!/usr/bin/env python3
"""snk_qnet_forward at the config-4 batch (65,536 samples), both precisions, device-timed; and the torch library paths."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from tools.torch_qnet import TorchQNet

S = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = S.SnakeGame(n)
out = env.alloc_outputs(obs="f32", mask=False)
for t in range(20):
    env.step_fused(act_idx=torch.randint(0, 3, (n,), device="cuda", dtype=torch.uint8), out=out)
obs = out["obs"]
layers = S.qnet.glorot_layers(0)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {"n": n}
want = TorchQNet(layers, obs.device, dtype=torch.float64)(obs.double())
for prec in ("f32", "bf16"):
    net = S.qnet.QNet(layers, env.device, precision=prec)
    ms = timed(lambda: net(obs))
    q = net(obs)
    res[prec] = {"ms": ms, "useful_tflops": 4870784.0 * n / (ms * 1e-3) / 1e12,
                 "max_err_of_maxQ": float(((q.double() - want).abs().max() / want.abs().max()).item())}
for name, kw in (("torch_fp32_no_tf32", dict(dtype=torch.float32)), ("torch_fp32_tf32", dict(dtype=torch.float32, allow_tf32=True)),
                 ("torch_bf16_channels_last", dict(dtype=torch.bfloat16, channels_last=True))):
    ref = TorchQNet(layers, obs.device, **kw)
    ms = timed(lambda: ref(obs), reps=5)
    res[name] = {"ms": ms, "max_err_of_maxQ": float(((ref(obs).double() - want).abs().max() / want.abs().max()).item())}
print(json.dumps(res))
