#!/usr/bin/env python3
"""BASELINE config 1: one env, 10,000 uniform random-action steps, a new game on loss — the per-step trace.

Writes one line per step:  t action_dir reward_bits(hex) done score board_hash(hex)
where board_hash = FNV-1a 64 over the 100 board cells (+1 each, column-major) AFTER the step (terminal board on a loss).
`bench_ref/trace_config1.jl` writes the same format from the unmodified Julia reference; the two files must be identical.
Source here: the CPU oracle (default) or the CUDA library (--cuda, needs a B200).
Actions: index = splitmix64(seed=42, env=0, t) % 3 into available_actions (tests/util.synth_actions).
"""
import argparse
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import synth_actions  # noqa: E402


def fnv1a(cells):
    h = 0xCBF29CE484222325
    for v in cells:
        h ^= (int(v) + 1) & 0xFF
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def trace_lines(steps=10000, cuda=False):
    acts = np.array([synth_actions(1, t)[0] for t in range(steps)], np.uint8)
    lines = []
    if cuda:
        import torch
        import __graft_entry__ as graft
        S = graft.load_package()
        env = S.SnakeGame(1, auto_reset=True)
        av_all = []
        out = env.alloc_outputs(obs="i8", mask=False, ep_stats=True)
        for t in range(steps):
            av = env.available_actions().cpu().numpy()[0]
            env.step_fused(act_idx=torch.from_numpy(acts[t:t + 1]).cuda(), out=out)
            board = out["obs"].cpu().numpy().reshape(200)[100:]
            lines.append("%d %d %08x %d %d %016x" % (t + 1, av[acts[t]], out["reward"].cpu().numpy().view(np.uint32)[0],
                                                    int(out["done"][0]), int(out["ep_score"][0]), fnv1a(board)))
    else:
        from oracle import oracle_lib as O
        ora = O.OracleBatch(1, auto_reset=True)
        for t in range(steps):
            av = ora.available_actions()[0]
            ref = ora.step(acts[t:t + 1], obs=("i8",))
            lines.append("%d %d %08x %d %d %016x" % (t + 1, av[acts[t]], ref["reward"].view(np.uint32)[0], int(ref["done"][0]),
                                                    int(ref["ep_score"][0]), fnv1a(ref["obs_i8"][0][100:])))
    return lines


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--cuda", action="store_true")
    ap.add_argument("--out", default="-")
    a = ap.parse_args()
    ls = trace_lines(a.steps, a.cuda)
    text = "\n".join(ls) + "\n"
    if a.out == "-":
        sys.stdout.write(text)
    else:
        open(a.out, "w").write(text)
    sys.stderr.write("sha256 %s\n" % hashlib.sha256(text.encode()).hexdigest())
