#!/usr/bin/env python3
"""Generate tests/golden/g8_qnet_on_gif_boards.json: the Float64 oracles of the Q-net forward (oracle/qnet_oracle.py) and of the
per-sample loss gradients (oracle/qgrad_oracle.py) evaluated on two-frame states taken from the REFERENCE'S OWN best game
(frames of trainer_gifs/very_long_double_training3.gif, already decoded into g2_boards_double3.npy by make_golden.py) with seeded
synthetic weights (the two-frame checkpoints are missing from the reference mount).  It pins nothing to the reference's
numerics — no Q-value is committed there — but it freezes the oracles (a CPU test recomputes them) and gives the GPU tests
reference-produced boards, including the terminal frame with the overwritten wall cell, as inputs.

  python tests/golden/make_golden_qnet.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from oracle import qgrad_oracle as QG  # noqa: E402
from oracle import qnet_oracle as QO  # noqa: E402

FRAMES = [1, 2, 30, 77, 150, 237, 238, 239]          # state = (board[t-1], board[t]); 238 = the terminal frame, 239 the padding copy
SEED = 7
BLOCKS = [("W1", 0, 288), ("b1", 288, 304), ("W2", 304, 4912), ("b2", 4912, 4944), ("W3", 4944, 78672), ("b3", 78672, 78736),
          ("W4", 78736, 181136), ("b4", 181136, 181200), ("W5", 181200, 181392), ("b5", 181392, 181395)]
PROBES = [5, 290, 1000, 4920, 40000, 78700, 100000, 181150, 181201, 181393]


def layers_for(S):
    layers = S.qnet.glorot_layers(seed=SEED)
    rng = np.random.default_rng(SEED + 100)
    for _, p in layers:
        if "b" in p:
            p["b"] = rng.normal(0, 0.05, p["b"].shape).astype(np.float32)
    return layers


def states(boards):
    """(B,2,10,10) in torch layout [n][frame][col][row] = Julia (10,10,2,B) column-major"""
    return np.ascontiguousarray(np.stack([np.stack([boards[t - 1].T, boards[t].T]) for t in FRAMES]).astype(np.float64))


def compute(S):
    boards = np.load(os.path.join(HERE, "g2_boards_double3.npy"))
    layers = layers_for(S)
    st = states(boards)
    q = QO.forward(layers, st.transpose(3, 2, 1, 0)).T                      # (B,3)
    act = q.argmax(1)
    y = np.array([q[i, act[i]] - 0.5 if i % 2 == 0 else q[i, act[i]] + 2.0 for i in range(len(FRAMES))])
    J, loss, _ = QG.per_sample_grads(layers, st, act, y)
    theta = S.bson_io.destructure(layers).astype(np.float32)
    return {"frames": FRAMES, "weights": {"init": "qnet.glorot_layers(seed=%d) + N(0, 0.05) biases from default_rng(%d)" % (SEED, SEED + 100),
                                          "theta_sha1": hashlib.sha1(theta.tobytes()).hexdigest()},
            "q": q.tolist(), "actions": act.tolist(), "targets": y.tolist(), "loss": loss.tolist(),
            "grad_block_norms": {name: np.linalg.norm(J[:, a:b], axis=1).tolist() for name, a, b in BLOCKS},
            "grad_probes": {str(i): J[:, i].tolist() for i in PROBES}}


if __name__ == "__main__":
    S = graft.load_package()
    out = compute(S)
    with open(os.path.join(HERE, "g8_qnet_on_gif_boards.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote g8_qnet_on_gif_boards.json:", len(out["frames"]), "states; Q[0] =", out["q"][0])
