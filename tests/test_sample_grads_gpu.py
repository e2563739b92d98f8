"""Per-sample gradients of the DQN loss (snk_qnet_sample_grads, csrc/qnet_grads.cu) vs the Float64 autograd oracle.

Reference: utils.jl:452-466 (Flux.huber_loss(q_net(s)[a], y)), structs.jl:127-139.  Stated tolerance (FP32 arithmetic
against Float64): every row within 2e-5 relative Frobenius error and 2e-5 of the row's largest entry elementwise; the
bf16 planes reproduce the FP32 row to 2^-16 relative per element; the Gram built from the planes within 1e-5.
Parity unpinned by the reference (it never forms per-sample gradients; Flux/Zygote unpinned)."""
import numpy as np
import pytest
import torch

from oracle import gram_oracle as GO
from oracle import qgrad_oracle as QG
from tests.util import pkg, synth_actions

pytestmark = pytest.mark.gpu


def _layers(seed):
    S = pkg()
    layers = S.qnet.glorot_layers(seed=seed)
    rng = np.random.default_rng(seed + 100)
    for _, p in layers:
        if "b" in p:
            p["b"] = rng.normal(0, 0.05, p["b"].shape).astype(np.float32)
    return layers


def _transitions(n, steps, B, seed):
    """B transitions out of a device replay ring filled by a real rollout"""
    S = pkg()
    env = S.SnakeGame(n, auto_reset=True)
    rb = S.ReplayBuffer(capacity=n * steps)
    out = env.alloc_outputs(obs=None, mask=True)
    for t in range(steps):
        env.step_fused(act_idx=torch.from_numpy(synth_actions(n, t, seed=seed)).cuda(), out=out, replay=rb)
    idx = torch.from_numpy(np.random.default_rng(seed).choice(n * steps, B, replace=False).astype(np.int64)).cuda()
    return rb.stack_exp(idx)


@pytest.mark.parametrize("B,seed", [(256, 1), (37, 2)])
def test_rows_match_float64_autograd(B, seed):
    S = pkg()
    layers = _layers(seed)
    batch = _transitions(512, 12, B, seed)
    net = S.qnet.QNet(layers, batch["states"].device, precision="f32")
    q = net(batch["states"])
    rng = np.random.default_rng(seed)
    # targets on both branches of the Huber loss: |q_sel - y| < 1 for most, > 1 for a quarter of the samples
    delta = np.where(rng.random(B) < 0.25, rng.choice([-1, 1], B) * rng.uniform(1.2, 3.0, B), rng.uniform(-0.9, 0.9, B))
    q_sel = q.gather(1, batch["actions"].long()[:, None])[:, 0].double().cpu().numpy()
    y = torch.from_numpy(q_sel - delta).cuda()
    plan = S.GramPlan(B, S.qnet.N_PARAMS, q.device)
    res = net.sample_grads(batch["states"], batch["actions"], y, planes=plan.planes(), want_J=True, want_loss=True)
    J = res["J"].cpu().numpy().astype(np.float64)
    want, loss, q64 = QG.per_sample_grads(layers, batch["states"].cpu().numpy(), batch["actions"].cpu().numpy(), y.cpu().numpy())
    assert np.abs(q.cpu().numpy() - q64).max() < 2e-5 * np.abs(q64).max()
    assert np.allclose(res["loss"].cpu().numpy(), loss, rtol=2e-5, atol=1e-7)
    norm = np.linalg.norm(want, axis=1)
    assert norm.min() > 0
    assert (np.linalg.norm(J - want, axis=1) / norm).max() < 2e-5
    assert (np.abs(J - want).max(1) / np.abs(want).max(1)).max() < 2e-5
    # the planes the Gram consumes: hi + lo == the FP32 row to 2^-16 per element, padding columns zero
    hi_p, lo_p, pitch = plan.planes()
    nb = B * pitch
    hi = plan.ws[:nb * 2].view(torch.bfloat16).view(B, pitch).float().cpu().numpy().astype(np.float64)
    lo = plan.ws[lo_p - hi_p:lo_p - hi_p + nb * 2].view(torch.bfloat16).view(B, pitch).float().cpu().numpy().astype(np.float64)
    P = S.qnet.N_PARAMS
    assert np.all(hi[:, P:] == 0) and np.all(lo[:, P:] == 0)
    rec = hi[:, :P] + lo[:, :P]
    assert np.all(np.abs(rec - J) <= 2.0 ** -16 * np.abs(J) + 1e-30)
    # Gram of the per-sample gradients straight from the planes (no pack pass) vs Float64
    G = plan.gram(terms=3).cpu().numpy().astype(np.float64)
    ref = GO.gram(want)
    assert np.linalg.norm(G - ref) / np.linalg.norm(ref) < 1e-5


def test_the_action_selects_the_output_row():
    """only row a of the W5 / b5 block of a sample's gradient is non-zero (q_sel = q_pred[a_i, i], utils.jl:454)"""
    S = pkg()
    layers = _layers(5)
    batch = _transitions(128, 6, 48, 5)
    net = S.qnet.QNet(layers, batch["states"].device, precision="f32")
    st, ac = batch["states"], batch["actions"]
    y0 = torch.full((48,), -7.0, dtype=torch.float64, device=st.device)
    probe = net.sample_grads(st, ac, y0, want_J=True)["J"]
    b5 = probe[:, -3:]                                     # d loss / d b5[a'] = g * [a' == a], g = 1 here (|d| > 1)
    assert torch.equal(b5, torch.nn.functional.one_hot(ac.long(), 3).float())
    W5 = probe[:, 181200:181392].reshape(48, 64, 3)        # (3,64) column-major = [c][a']
    other = torch.ones(48, 3, dtype=torch.bool, device=st.device)
    other.scatter_(1, ac.long()[:, None], False)
    assert (W5.permute(0, 2, 1)[other] == 0).all()
