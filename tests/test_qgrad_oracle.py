"""The per-sample-gradient oracle (oracle/qgrad_oracle.py, torch autograd in Float64) against central finite differences of the
independent numpy restatement of the network (oracle/qnet_oracle.py): pins the Flux.destructure order and the conv-weight
flips of the oracle the GPU kernel is tested with."""
import numpy as np

from oracle import qgrad_oracle as QG
from oracle import qnet_oracle as QO
from tests.util import pkg


def test_autograd_rows_match_finite_differences_of_the_numpy_network():
    S = pkg()
    layers = S.qnet.glorot_layers(1)
    rng = np.random.default_rng(0)
    for _, p in layers:
        if "b" in p:
            p["b"] = rng.normal(0, 0.05, p["b"].shape).astype(np.float32)
    st = rng.integers(-1, 3, (2, 2, 10, 10)).astype(np.float64)
    acts, ys = [0, 2], [0.1, 5.0]                         # quadratic and linear branch of the Huber loss
    J, loss, q = QG.per_sample_grads(layers, st, acts, ys)
    assert J.shape == (2, 181395)
    assert np.allclose(q, QO.forward(layers, st.transpose(3, 2, 1, 0)).T, rtol=1e-12, atol=1e-12)
    th = S.bson_io.destructure(layers).astype(np.float64)

    def loss_of(theta, i):
        ls, off = [], 0
        for k, p in layers:
            if k in ("conv", "dense"):
                W = theta[off:off + p["W"].size].reshape(p["W"].shape, order="F"); off += p["W"].size
                b = theta[off:off + p["b"].size]; off += p["b"].size
                d = dict(p); d["W"], d["b"] = W, b
                ls.append((k, d))
            else:
                ls.append((k, p))
        d = QO.forward(ls, st[i:i + 1].transpose(3, 2, 1, 0))[acts[i], 0] - ys[i]
        return 0.5 * d * d if abs(d) < 1 else abs(d) - 0.5

    # one entry in every parameter block: W1 b1 W2 b2 W3 b3 W4 b4 W5 b5
    for i in range(2):
        for idx in (5, 290, 1000, 4920, 40000, 78700, 100000, 181150, 181200 + acts[i], 181392 + acts[i]):
            e = np.zeros_like(th); e[idx] = 1e-6
            fd = (loss_of(th + e, i) - loss_of(th - e, i)) / 2e-6
            assert abs(fd - J[i, idx]) < 1e-7 * max(1.0, abs(fd)), (i, idx, fd, J[i, idx])
