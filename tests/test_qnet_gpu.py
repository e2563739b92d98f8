"""Native Q-net forward (tcgen05 implicit-GEMM convolutions) vs the Float64 Flux-semantics oracle."""
import numpy as np
import pytest
import torch

from oracle import qnet_oracle as QO
from tests.util import pkg, synth_actions

pytestmark = pytest.mark.gpu


def _layers(seed, bias=True):
    S = pkg()
    layers = S.qnet.glorot_layers(seed=seed)
    if bias:
        rng = np.random.default_rng(seed + 100)
        for _, p in layers:
            if "b" in p:
                p["b"] = rng.normal(0, 0.05, p["b"].shape).astype(np.float32)
    return layers


def _real_obs(n, steps=12):
    """boards the env really produces (walls, snake, food), after a few random steps"""
    S = pkg()
    env = S.SnakeGame(n, auto_reset=True)
    out = env.alloc_outputs(obs="f32", mask=False)
    for t in range(steps):
        env.step_fused(act_idx=torch.from_numpy(synth_actions(n, t, seed=3)).cuda(), out=out)
    return out["obs"].clone()


@pytest.mark.parametrize("engine", [17, 16, 12])
@pytest.mark.parametrize("n", [1, 12, 13, 16, 17, 100, 257, 5000])
def test_native_forward_matches_oracle(n, engine, monkeypatch):
    """both conv engines (snk_qnet_create reads SNK_QNET_ENGINE; 17 = default) at ragged sizes around their
    samples-per-iteration (12 / 16) and one size that gives every CTA several iterations"""
    monkeypatch.setenv("SNK_QNET_ENGINE", str(engine))
    S = pkg()
    layers = _layers(seed=n)
    obs = _real_obs(n)
    net = S.qnet.QNet(layers, obs.device, backend="native")
    q = net(obs).cpu().numpy()
    m = min(n, 300)
    state_julia = obs[:m].cpu().numpy().astype(np.float64).transpose(3, 2, 1, 0)      # (10,10,2,m)
    want = QO.forward(layers, state_julia).T                                           # (m,3)
    scale = np.abs(want).max()
    err = np.abs(q[:m] - want).max() / scale
    # bf16 operands (2^-9 per rounding) through 5 layers with fp32 accumulation: stated tolerance 1.5e-2 of max|Q|
    assert err < 1.5e-2, err
    ref32 = S.qnet.QNet(layers, obs.device, backend="torch")(obs).cpu().numpy()
    assert np.abs(q - ref32).max() / scale < 1.5e-2
    # argmax agrees wherever the fp32 margin is clear
    srt = np.sort(ref32, axis=1)
    clear = (srt[:, 2] - srt[:, 1]) > 4e-2 * scale
    assert clear.mean() > 0.3
    assert np.array_equal(q[clear].argmax(1), ref32[clear].argmax(1))


def test_native_forward_random_inputs_and_linearity_in_last_layer():
    """arbitrary (non-board) inputs; and a property: Q is affine in the last layer's bias."""
    S = pkg()
    layers = _layers(seed=7)
    rng = np.random.default_rng(0)
    obs = torch.from_numpy(rng.integers(-1, 3, (777, 2, 10, 10)).astype(np.float32)).cuda()
    want = QO.forward(layers, obs.cpu().numpy().astype(np.float64).transpose(3, 2, 1, 0)).T
    q = S.qnet.QNet(layers, obs.device, backend="native")(obs).cpu().numpy()
    assert np.abs(q - want).max() / np.abs(want).max() < 1.5e-2
    layers2 = [(k, dict(p)) for k, p in layers]
    layers2[-1][1]["b"] = layers[-1][1]["b"] + np.array([1.0, -2.0, 0.5], np.float32)
    q2 = S.qnet.QNet(layers2, obs.device, backend="native")(obs).cpu().numpy()
    assert np.allclose(q2 - q, np.array([1.0, -2.0, 0.5]), atol=1e-5)


def test_native_forward_at_the_config4_batch():
    """65,536 samples = 28 iterations per CTA: every hand-over between iterations (accumulator reuse, the borrowed conv2 buffers,
    the parked halves) is exercised many times; compared with the Float32 library path on every sample, twice (a second call
    must give the same bits: no state leaks from one launch into the next)."""
    S = pkg()
    layers = _layers(seed=21)
    obs = _real_obs(65536, steps=20)
    net = S.qnet.QNet(layers, obs.device, backend="native")
    q1 = net(obs).clone()
    q2 = net(obs)
    assert torch.equal(q1, q2)
    want = S.qnet.QNet(layers, obs.device, backend="torch")(obs)
    scale = want.abs().max()
    assert ((q1 - want).abs().max() / scale).item() < 1.5e-2
    assert torch.isfinite(q1).all()


def test_one_handle_many_batch_sizes():
    """the same snk_qnet handle called with growing and shrinking N (the conv3 activation buffer is re-allocated on growth)"""
    S = pkg()
    layers = _layers(seed=11)
    net = S.qnet.QNet(layers, torch.device("cuda", 0), backend="native")
    ref = S.qnet.QNet(layers, torch.device("cuda", 0), backend="torch")
    for n in (100, 5000, 17, 2368, 2369):          # 2368 = 148 SMs x 16 samples: exactly one iteration per CTA, then one more
        obs = _real_obs(n, steps=5)
        q, want = net(obs).cpu().numpy(), ref(obs).cpu().numpy()
        assert q.shape == (n, 3)
        assert np.abs(q - want).max() / np.abs(want).max() < 1.5e-2, n


def test_rollout_with_native_qnet_runs_config4_shape():
    S = pkg()
    n = 4096
    env = S.SnakeGame(n, auto_reset=True)
    rb = S.ReplayBuffer(capacity=50000)
    net = S.qnet.QNet(_layers(seed=1), env.device, backend="native")
    ro = S.rollout.Rollout(env, net, net, rb, epsilon=0.05)
    for _ in range(15):
        res = ro.step()
    assert len(rb) == 50000 and res["target"].dtype == torch.float64 and res["target"].shape == (n,)
    assert torch.isfinite(res["target"]).all() and env.count_errors() == 0
    batch = rb.sample()
    assert batch["states"].shape == (64, 2, 10, 10)
