"""Native Q-net forward (tcgen05 implicit-GEMM convolutions) vs the Float64 Flux-semantics oracle.

Two precisions (snk_qnet_create): "f32" = Float32-faithful (fp16 hi/lo split operands, four products) — the mode the
reference's Float32 network needs; "bf16" = fast mode.  Stated tolerances, both relative to max|Q| of the batch:
    f32 : 2e-5 against Float64 (measured ~2e-6); argmax identical on EVERY sample whose Float64 top-2 gap exceeds 1e-4
    bf16: 1.5e-2; argmax only where the gap exceeds 4e-2
Parity is unpinned by the reference (no Q-values or two-frame weights are committed): the oracle is the numpy restatement
of Flux/NNlib semantics (oracle/qnet_oracle.py), cross-checked by torch's Float64 convolution (tools/torch_qnet.py).
"""
import numpy as np
import pytest
import torch

from oracle import qnet_oracle as QO
from tests.util import pkg, synth_actions
from tools.torch_qnet import TorchQNet

pytestmark = pytest.mark.gpu

TOL = {"f32": 2e-5, "bf16": 1.5e-2}
GAP = {"f32": 1e-4, "bf16": 4e-2}


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


def _check_against(q, want, precision, what):
    """q, want: (n,3) arrays; want in Float64.  Error bound + argmax agreement on every sample with a clear gap."""
    scale = np.abs(want).max()
    err = np.abs(q - want).max() / scale
    assert err < TOL[precision], (what, precision, err)
    srt = np.sort(want, axis=1)
    clear = (srt[:, 2] - srt[:, 1]) > GAP[precision] * scale
    assert np.array_equal(q[clear].argmax(1), want[clear].argmax(1)), (what, precision)
    return err, clear.mean()


@pytest.mark.parametrize("precision", ["f32", "bf16"])
@pytest.mark.parametrize("n", [1, 7, 8, 9, 16, 17, 100, 257, 5000])
def test_native_forward_matches_oracle(n, precision):
    """ragged sizes around the samples-per-iteration (8 in f32 mode, 16 in bf16 mode) and one size that gives every CTA
    several iterations; the first 300 samples against the numpy oracle, all of them against torch Float64"""
    S = pkg()
    layers = _layers(seed=n)
    obs = _real_obs(n)
    net = S.qnet.QNet(layers, obs.device, precision=precision)
    q = net(obs).cpu().numpy().astype(np.float64)
    m = min(n, 300)
    state_julia = obs[:m].cpu().numpy().astype(np.float64).transpose(3, 2, 1, 0)      # (10,10,2,m)
    want = QO.forward(layers, state_julia).T                                           # (m,3)
    _check_against(q[:m], want, precision, "numpy oracle")
    want_all = TorchQNet(layers, obs.device, dtype=torch.float64)(obs.double()).cpu().numpy()
    assert np.abs(want_all[:m] - want).max() < 1e-9 * max(1.0, np.abs(want).max())     # the two Float64 evaluations agree
    err, clear = _check_against(q, want_all, precision, "torch float64")
    if precision == "f32":
        assert not net.overflow()
        if n >= 100:
            assert clear > 0.9        # the gap criterion covers almost every sample


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_native_forward_random_inputs_and_linearity_in_last_layer(precision):
    """arbitrary (non-board) inputs; and a property: Q is affine in the last layer's bias."""
    S = pkg()
    layers = _layers(seed=7)
    rng = np.random.default_rng(0)
    obs = torch.from_numpy(rng.integers(-1, 3, (777, 2, 10, 10)).astype(np.float32)).cuda()
    want = QO.forward(layers, obs.cpu().numpy().astype(np.float64).transpose(3, 2, 1, 0)).T
    q = S.qnet.QNet(layers, obs.device, precision=precision)(obs).cpu().numpy()
    assert np.abs(q - want).max() / np.abs(want).max() < TOL[precision]
    layers2 = [(k, dict(p)) for k, p in layers]
    layers2[-1][1]["b"] = layers[-1][1]["b"] + np.array([1.0, -2.0, 0.5], np.float32)
    q2 = S.qnet.QNet(layers2, obs.device, precision=precision)(obs).cpu().numpy()
    assert np.allclose(q2 - q, np.array([1.0, -2.0, 0.5]), atol=1e-5)


def test_f32_mode_with_non_integer_inputs_and_small_weights():
    """Float32 inputs that are not board values (conv1 runs in plain FP32) and a net whose weights are 1/64 of Glorot's
    (low halves of the split deep in fp16's subnormal range: the absolute error stays 2^-25 per weight)"""
    S = pkg()
    layers = _layers(seed=5)
    for _, p in layers:
        if "W" in p:
            p["W"] = (p["W"] / 64).astype(np.float32)
    rng = np.random.default_rng(1)
    obs = torch.from_numpy(rng.normal(0, 1.5, (333, 2, 10, 10)).astype(np.float32)).cuda()
    want = TorchQNet(layers, obs.device, dtype=torch.float64)(obs.double()).cpu().numpy()
    q = S.qnet.QNet(layers, obs.device, precision="f32")(obs).cpu().numpy()
    assert np.abs(q - want).max() / np.abs(want).max() < 1e-4


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_native_forward_at_the_config4_batch(precision):
    """65,536 samples = many iterations per CTA: every hand-over between iterations (accumulator reuse, weight ring,
    parked halves) is exercised many times; EVERY sample against torch Float64, twice (a second call must give the
    same bits: no state leaks from one launch into the next)."""
    S = pkg()
    layers = _layers(seed=21)
    obs = _real_obs(65536, steps=20)
    net = S.qnet.QNet(layers, obs.device, precision=precision)
    q1 = net(obs).clone()
    q2 = net(obs)
    assert torch.equal(q1, q2)
    assert torch.isfinite(q1).all()
    want = TorchQNet(layers, obs.device, dtype=torch.float64)(obs.double()).cpu().numpy()
    err, clear = _check_against(q1.cpu().numpy().astype(np.float64), want, precision, "config-4 batch")
    print("config-4 batch, %s: max err %.3g of max|Q|, %.2f%% of samples above the gap criterion" % (precision, err, 100 * clear))
    if precision == "f32":
        # against the Float32 library evaluation (cuDNN, TF32 off) the split mode must be at least as close to Float64
        lib32 = TorchQNet(layers, obs.device, dtype=torch.float32)(obs).cpu().numpy().astype(np.float64)
        err_lib = np.abs(lib32 - want).max() / np.abs(want).max()
        print("Float32 library path vs Float64: %.3g" % err_lib)
        assert err < 20 * max(err_lib, 1e-7)


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_one_handle_many_batch_sizes(precision):
    """the same snk_qnet handle called with growing and shrinking N (the conv3 activation buffer is re-allocated on growth)"""
    S = pkg()
    layers = _layers(seed=11)
    net = S.qnet.QNet(layers, torch.device("cuda", 0), precision=precision)
    ref = TorchQNet(layers, torch.device("cuda", 0), dtype=torch.float64)
    for n in (100, 5000, 17, 1184, 1185, 2368, 2369):   # 148 SMs x 8 / x 16 samples: exactly one iteration per CTA, then one more
        obs = _real_obs(n, steps=5)
        q, want = net(obs).cpu().numpy(), ref(obs.double()).cpu().numpy()
        assert q.shape == (n, 3)
        assert np.abs(q - want).max() / np.abs(want).max() < TOL[precision], n


def test_f32_mode_flags_activations_outside_the_fp16_range():
    S = pkg()
    layers = _layers(seed=3)
    obs = _real_obs(64, steps=3)
    ok = S.qnet.QNet(layers, obs.device, precision="f32")
    ok(obs)
    assert not ok.overflow()
    net = S.qnet.QNet(layers, obs.device, precision="f32")
    net(torch.full((64, 2, 10, 10), 1.0e6, device=obs.device))            # conv1 outputs ~5e5 > 65504
    assert net.overflow()
    assert not net.overflow()                                              # reading clears the flag
    net(obs)
    assert not net.overflow()
    huge = [(k, dict(p)) for k, p in layers]
    huge[1][1]["W"] = (layers[1][1]["W"] * 1e7).astype(np.float32)
    with pytest.raises(S.SnakeB200Error):                                  # a weight outside the fp16 range is refused at create
        S.qnet.QNet(huge, obs.device, precision="f32")
    with pytest.raises(ValueError):
        S.qnet.QNet(layers, obs.device, precision="tf32")


def test_device_of_the_caller_is_left_alone():
    """every ABI entry runs on its handle's device and restores the caller's current device"""
    S = pkg()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(0)
    env = S.SnakeGame(64, device=1)
    env.step(torch.zeros(64, dtype=torch.uint8, device="cuda:1"))
    assert torch.cuda.current_device() == 0
