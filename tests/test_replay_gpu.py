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
