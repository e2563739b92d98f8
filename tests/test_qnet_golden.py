"""Q-net forward and per-sample gradients on boards of the reference's own best game (GIF frames, tests/golden/g2_*) against
the committed Float64 oracle values (tests/golden/g8_qnet_on_gif_boards.json, made by tests/golden/make_golden_qnet.py):
the CPU test freezes the oracles, the GPU tests check the kernels on reference-produced inputs (incl. the terminal frame with the
overwritten wall cell).  Weights are seeded synthetic ones: the reference holds no Q-value to pin (parity unpinned)."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.golden import make_golden_qnet as MG
from tests.util import ROOT, pkg

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "g8_qnet_on_gif_boards.json")))


def test_oracles_reproduce_the_committed_values():
    S = pkg()
    got = MG.compute(S)
    assert got["weights"]["theta_sha1"] == GOLD["weights"]["theta_sha1"]
    assert got["frames"] == GOLD["frames"] and got["actions"] == GOLD["actions"]
    assert np.allclose(got["q"], GOLD["q"], rtol=1e-12, atol=1e-14)
    assert np.allclose(got["loss"], GOLD["loss"], rtol=1e-12, atol=1e-14)
    for k in GOLD["grad_block_norms"]:
        assert np.allclose(got["grad_block_norms"][k], GOLD["grad_block_norms"][k], rtol=1e-10, atol=1e-14), k
    for k in GOLD["grad_probes"]:
        assert np.allclose(got["grad_probes"][k], GOLD["grad_probes"][k], rtol=1e-10, atol=1e-14), k


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_native_forward_on_gif_boards(precision):
    import torch
    S = pkg()
    boards = np.load(os.path.join(ROOT, "tests", "golden", "g2_boards_double3.npy"))
    obs = torch.from_numpy(MG.states(boards).astype(np.float32)).cuda()
    net = S.qnet.QNet(MG.layers_for(S), obs.device, precision=precision)
    assert hashlib.sha1(net.theta.tobytes()).hexdigest() == GOLD["weights"]["theta_sha1"]
    q = net(obs).cpu().numpy().astype(np.float64)
    want = np.array(GOLD["q"])
    tol = 2e-5 if precision == "f32" else 1.5e-2
    assert np.abs(q - want).max() / np.abs(want).max() < tol
    if precision == "f32":
        assert q.argmax(1).tolist() == GOLD["actions"]


@pytest.mark.gpu
def test_sample_grads_on_gif_boards():
    import torch
    S = pkg()
    boards = np.load(os.path.join(ROOT, "tests", "golden", "g2_boards_double3.npy"))
    obs = torch.from_numpy(MG.states(boards).astype(np.float32)).cuda()
    net = S.qnet.QNet(MG.layers_for(S), obs.device, precision="f32")
    res = net.sample_grads(obs, torch.tensor(GOLD["actions"], dtype=torch.uint8, device="cuda"),
                           torch.tensor(GOLD["targets"], dtype=torch.float64, device="cuda"), want_J=True, want_loss=True)
    J = res["J"].cpu().numpy().astype(np.float64)
    assert np.allclose(res["loss"].cpu().numpy(), GOLD["loss"], rtol=3e-5, atol=1e-7)
    for name, a, b in MG.BLOCKS:
        got, want = np.linalg.norm(J[:, a:b], axis=1), np.array(GOLD["grad_block_norms"][name])
        assert np.all(np.abs(got - want) <= 3e-5 * want + 1e-9), name
    rownorm = np.linalg.norm(J, axis=1)
    for k, want in GOLD["grad_probes"].items():
        assert np.all(np.abs(J[:, int(k)] - np.array(want)) <= 2e-5 * rownorm), k
