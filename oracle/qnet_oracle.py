"""Oracle (test infrastructure): the reference Q-network forward in Float64 numpy, restating Flux/NNlib.

structs.jl:127-139.  Flux/NNlib are third-party dependencies with unpinned versions (no Manifest.toml in the
reference); the published semantics restated here: Conv is a TRUE convolution (kernel flipped) over WHCN
arrays, y[i1,i2,o] = b[o] + sum_{a1,a2,c} W[a1,a2,c,o] * xpad[i1 + K1 - a1, i2 + K2 - a2, c] (1-based a),
Flux.flatten is a column-major reshape, Dense is W*x + b.  No reference test pins Q-values: parity unpinned.
"""
import numpy as np


def conv_true(x, W, b, pad):
    """x: (d1, d2, C, N) Float64, W: (K1, K2, C, O)."""
    K1, K2, C, O = W.shape
    d1, d2, _, N = x.shape
    xp = np.zeros((d1 + 2 * pad, d2 + 2 * pad, C, N))
    xp[pad:pad + d1, pad:pad + d2] = x
    o1, o2 = d1 + 2 * pad - K1 + 1, d2 + 2 * pad - K2 + 1
    y = np.zeros((o1, o2, O, N))
    for a1 in range(K1):
        for a2 in range(K2):
            patch = xp[K1 - 1 - a1:K1 - 1 - a1 + o1, K2 - 1 - a2:K2 - 1 - a2 + o2]      # (o1,o2,C,N)
            y += np.einsum("ijcn,co->ijon", patch, W[a1, a2])
    return y + b[None, None, :, None]


def forward(layers, state):
    """state: (10,10,C,N) (Julia layout, any real dtype).  Returns Q (3, N) Float64."""
    x = np.asarray(state, dtype=np.float64)
    for kind, p in layers:
        if kind == "conv":
            x = np.maximum(conv_true(x, p["W"].astype(np.float64), p["b"].astype(np.float64), int(p["pad"][0])), 0.0)
        elif kind == "flatten":
            x = x.reshape(-1, x.shape[-1], order="F")
        elif kind == "dense":
            x = p["W"].astype(np.float64) @ x + p["b"].astype(np.float64)[:, None]
            if p["W"].shape[0] != 3:
                x = np.maximum(x, 0.0)
    return x
