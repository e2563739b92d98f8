"""Device replay ring vs the reference's ReplayBuffer semantics (store! order, position quirk, stack_exp layout)."""
import numpy as np
import pytest
import torch

from oracle import oracle_lib as O
from oracle.replay_oracle import ReplayOracle
from tests.util import bits, pkg, synth_actions

pytestmark = pytest.mark.gpu


def _run(n, cap, steps, seed):
    S = pkg()
    env = S.SnakeGame(n, auto_reset=True)
    ora = O.OracleBatch(n, auto_reset=True)
    rb = S.ReplayBuffer(capacity=cap)
    ro = ReplayOracle(capacity=cap)
    out = env.alloc_outputs(obs="i8", mask=True, ep_stats=True)
    for t in range(steps):
        act = synth_actions(n, t, seed=seed)
        state = ora.state("i8")
        ref = ora.step(act, obs=("i8",))
        for i in range(n):                              # sequential store!, env order
            ro.store({"state": state[i], "next_state": ref["obs_i8"][i], "action_idx": act[i],
                      "reward": ref["reward"][i], "done": ref["done"][i], "mask": ref["mask"][i]})
        env.step_fused(act_idx=torch.from_numpy(act).cuda(), out=out, replay=rb)
        assert len(rb) == len(ro) and rb.position == ro.position, t
        assert rb.isfull() == ro.isfull() and rb.isready() == ro.isready()
    return S, rb, ro


def _cmp(rb, ro, idx):
    got = rb.stack_exp(torch.from_numpy(idx).cuda())
    want = ro.stack_exp(idx)
    assert not rb.bad_index()
    for k in ("states", "next_states"):
        assert np.array_equal(got[k].cpu().numpy().reshape(len(idx), 200), want[k]), k
    assert np.array_equal(got["actions"].cpu().numpy(), want["actions"])
    assert np.array_equal(bits(got["rewards"].cpu().numpy()), bits(want["rewards"]))
    assert np.array_equal(got["dones"].cpu().numpy(), want["dones"])
    assert np.array_equal(got["mask"].cpu().numpy(), want["mask"])


@pytest.mark.parametrize("n,cap,steps", [(64, 1000, 40), (100, 777, 30), (1000, 640, 7), (3, 64, 100)])
def test_store_order_and_stack_exp_match_reference_semantics(n, cap, steps):
    """incl. wrap-around, the position-only-advances-when-full quirk, and more envs per step than slots"""
    S, rb, ro = _run(n, cap, steps, seed=5)
    L = len(ro)
    _cmp(rb, ro, np.arange(L, dtype=np.int64))
    rng = np.random.default_rng(0)
    _cmp(rb, ro, rng.integers(0, L, 333).astype(np.int64))


def test_sample_without_replacement_and_empty():
    S, rb, ro = _run(256, 5000, 12, seed=9)
    assert len(rb) == 3072
    for _ in range(5):
        idx = rb.sample_indices().cpu().numpy()
        assert idx.shape == (64,) and len(set(idx.tolist())) == 64 and idx.min() >= 0 and idx.max() < 3072
    big = rb.sample_indices(3072).cpu().numpy()
    assert sorted(big.tolist()) == list(range(3072))            # a permutation of the filled part
    a, b = rb.sample_indices(512).cpu().numpy(), rb.sample_indices(512).cpu().numpy()
    assert not np.array_equal(a, b)
    hist = np.bincount(np.concatenate([rb.sample_indices(64).cpu().numpy() for _ in range(400)]) // 512, minlength=6)
    assert hist.min() > 0.8 * hist.mean()                       # roughly uniform
    batch = rb.sample()
    assert batch["states"].shape == (64, 2, 10, 10) and batch["mask"].shape == (64, 3)
    with pytest.raises(S.SnakeB200Error):
        rb.sample_indices(4000)
    rb.empty_buffer()
    assert len(rb) == 0 and rb.position == 1 and not rb.isready()
    with pytest.raises(ValueError):
        S.ReplayBuffer(capacity=10)


def test_transition_chain_state_is_previous_next_state():
    """state of the transition stored at step t+1 == next_state stored at step t unless the env reset."""
    S = pkg()
    n, cap = 512, 512 * 6
    env = S.SnakeGame(n, auto_reset=True)
    rb = S.ReplayBuffer(capacity=cap)
    out = env.alloc_outputs(obs=None, mask=False)
    for t in range(6):
        env.step_fused(act_idx=torch.from_numpy(synth_actions(n, t, seed=2)).cuda(), out=out, replay=rb)
    allb = rb.stack_exp(torch.arange(cap, device="cuda"))
    s = allb["states"].view(6, n, 2, 10, 10)
    ns = allb["next_states"].view(6, n, 2, 10, 10)
    d = allb["dones"].view(6, n)
    for t in range(5):
        live = d[t] == 0
        assert torch.equal(s[t + 1][live], ns[t][live])
        assert torch.equal(s[t + 1][~live][:, 0], s[t + 1][~live][:, 1])      # fresh game: (init, init)
    assert torch.equal(s[:, :, 1], ns[:, :, 0])                                # shared middle board


def test_host_buffer_store_path_equals_device_path():
    """snk_step_fused_store_host (pinned host buffers, 2-bit packed observations, chunked + pipelined copies: the call bench.py's
    e2e leg times) against snk_step_fused_store on device buffers: same outputs, same ring contents, several steps in a row
    without a sync in between two calls (the staging buffers are guarded by the copy-done event)."""
    S = pkg()
    n, cap, steps = 200_000, 50_000, 6          # several chunks per step, more envs per step than ring slots
    env_d, env_h = S.SnakeGame(n, auto_reset=True), S.SnakeGame(n, auto_reset=True)
    rb_d, rb_h = S.ReplayBuffer(capacity=cap), S.ReplayBuffer(capacity=cap)
    out = env_d.alloc_outputs(obs="packed2", mask=True, act=True)
    host = {"obs_fmt": "packed2", "q": S.pinned_empty((n, 3), torch.float32), "u": S.pinned_empty((n,), torch.float32),
            "ridx": S.pinned_empty((n,), torch.uint8), "act_idx": S.pinned_empty((n,), torch.uint8),
            "reward": S.pinned_empty((n,), torch.float32), "done": S.pinned_empty((n,), torch.uint8),
            "obs": S.pinned_empty((n, 50), torch.uint8), "mask": S.pinned_empty((n, 3), torch.uint8)}
    g = torch.Generator().manual_seed(11)
    for t in range(steps):
        env_h.sync()                                   # the host inputs are about to be overwritten
        host["q"].copy_(torch.rand(n, 3, generator=g) * 2 - 1)
        host["u"].copy_(torch.rand(n, generator=g))
        host["ridx"].copy_(torch.randint(0, 3, (n,), generator=g, dtype=torch.uint8))
        env_d.step_fused(q=host["q"].cuda(), eps=0.2, u=host["u"].cuda(), ridx=host["ridx"].cuda(), out=out, replay=rb_d)
        env_h.step_fused_host(host, q=True, eps=0.2, replay=rb_h)
        if t % 2 == 1:                                 # every other step: enqueue the next call's getters right behind it
            assert torch.equal(env_h.score_host(), env_d.score.cpu())
        env_h.sync()
        for k in ("act_idx", "reward", "done", "obs", "mask"):
            assert torch.equal(out[k].cpu(), host[k]), (k, t)
        assert len(rb_h) == len(rb_d) and rb_h.position == rb_d.position
    idx = torch.arange(cap, device="cuda")
    a, b = rb_d.stack_exp(idx, ep_stats=True), rb_h.stack_exp(idx, ep_stats=True)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert env_h.count_errors() == 0


def test_host_store_path_with_bit_records():
    """the same host entry with SNK_OBS_BITS: ONE 24-byte record per env comes down (no separate reward / done / mask / action
    arrays) — the format bench.py's e2e leg uses — and decodes to what the device path returns in the int8 format"""
    S = pkg()
    n, cap, steps = 150_001, 40_000, 5
    env_d, env_h = S.SnakeGame(n, auto_reset=True), S.SnakeGame(n, auto_reset=True)
    rb_d, rb_h = S.ReplayBuffer(capacity=cap), S.ReplayBuffer(capacity=cap)
    out = env_d.alloc_outputs(obs="i8", mask=True, act=True)
    host = {"obs_fmt": "bits", "q": S.pinned_empty((n, 3), torch.float32), "u": S.pinned_empty((n,), torch.float32),
            "ridx": S.pinned_empty((n,), torch.uint8), "obs": S.pinned_empty((n, 24), torch.uint8)}
    g = torch.Generator().manual_seed(5)
    for t in range(steps):
        env_h.sync()
        host["q"].copy_(torch.rand(n, 3, generator=g) * 2 - 1)
        host["u"].copy_(torch.rand(n, generator=g))
        host["ridx"].copy_(torch.randint(0, 3, (n,), generator=g, dtype=torch.uint8))
        env_d.step_fused(q=host["q"].cuda(), eps=0.2, u=host["u"].cuda(), ridx=host["ridx"].cuda(), out=out, replay=rb_d)
        env_h.step_fused_host(host, q=True, eps=0.2, replay=rb_h)
        env_h.sync()
        d = S.unpack_bits(host["obs"])
        assert torch.equal(d["state"], out["obs"].cpu()), t
        assert torch.equal(d["reward"], out["reward"].cpu()) and torch.equal(d["done"], out["done"].cpu()), t
        assert torch.equal(d["mask"], out["mask"].cpu()) and torch.equal(d["action"], out["act_idx"].cpu()), t
    idx = torch.arange(cap, device="cuda")
    a, b = rb_d.stack_exp(idx), rb_h.stack_exp(idx)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert env_h.count_errors() == 0

