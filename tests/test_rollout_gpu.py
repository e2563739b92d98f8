"""Differential test of the batched play_episode (rollout.Rollout.step) against the oracle driven with the same Q.

Reference: utils.jl:198-259 (play_episode: epsilon_greedy -> step! -> virtual_step, Experience assembly),
utils.jl:267-277 (store! order), utils.jl:448-451 (masked max-Q target).  The GPU side runs the native Float32-faithful
Q-net + the fused kernel + the device replay ring; the oracle side steps oracle/snake_oracle.c envs and evaluates the
network in Float64 (torch, cross-checked against oracle/qnet_oracle.py in test_qnet_gpu / test_bson_qnet).

Both sides get the same injected draws (u, ridx).  The oracle decides the action from ITS Float64 Q; wherever the Float64
top-2 gap exceeds 1e-4 of max|Q| the GPU's action must be identical — below that either maximum is accepted and the oracle
env follows the GPU's choice so that the two stay in lockstep (counted; must be rare).
"""
import numpy as np
import pytest
import torch

from oracle import oracle_lib as O
from oracle.replay_oracle import ReplayOracle
from tests.util import bits, pkg
from tools.torch_qnet import TorchQNet

pytestmark = pytest.mark.gpu


def _layers(seed):
    S = pkg()
    layers = S.qnet.glorot_layers(seed=seed)
    rng = np.random.default_rng(seed + 100)
    for _, p in layers:
        if "b" in p:
            p["b"] = rng.normal(0, 0.05, p["b"].shape).astype(np.float32)
    return layers


@pytest.mark.parametrize("n,steps,cap,eps", [(4096, 200, 50000, 0.05), (333, 60, 1000, 0.3)])
def test_rollout_step_matches_oracle_driven_episode(n, steps, cap, eps):
    S = pkg()
    dev = torch.device("cuda", 0)
    q_layers, t_layers = _layers(1), _layers(2)
    env = S.SnakeGame(n, auto_reset=True)
    rb = S.ReplayBuffer(capacity=cap)
    ro = S.rollout.Rollout(env, S.qnet.QNet(q_layers, dev, "f32"), S.qnet.QNet(t_layers, dev, "f32"), rb, epsilon=eps)
    q64, t64 = TorchQNet(q_layers, dev, dtype=torch.float64), TorchQNet(t_layers, dev, dtype=torch.float64)
    ora = O.OracleBatch(n, auto_reset=True)
    rpo = ReplayOracle(capacity=cap)
    rng = np.random.default_rng(7)
    near_ties = 0
    for t in range(steps):
        u = rng.random(n, dtype=np.float32)
        ridx = rng.integers(0, 3, n).astype(np.uint8)
        # ---- oracle side: state -> Float64 Q -> epsilon_greedy (utils.jl:153-172)
        state = ora.state("i8")                                                   # (n,200) int8, Julia layout per env
        s_t = torch.from_numpy(state.reshape(n, 2, 10, 10).astype(np.float64)).to(dev)
        qo = q64(s_t).cpu().numpy()                                               # (n,3) Float64
        greedy = qo.argmax(1)
        want_act = np.where(u < np.float32(eps), ridx, greedy).astype(np.uint8)
        # ---- GPU side
        res = ro.step(u=torch.from_numpy(u).to(dev), ridx=torch.from_numpy(ridx).to(dev))
        act = res["act_idx"].cpu().numpy()
        scale = np.abs(qo).max()
        srt = np.sort(qo, axis=1)
        clear = ((srt[:, 2] - srt[:, 1]) > 1e-4 * scale) | (u < np.float32(eps))
        assert np.array_equal(act[clear], want_act[clear]), t
        diff = act != want_act
        if diff.any():                                                            # near tie: the GPU must still have picked a maximum
            i = np.nonzero(diff)[0]
            assert np.all(qo[i, act[i]] >= srt[i, 2] - 1e-4 * scale), t
            near_ties += len(i)
        assert np.abs(res["q"].cpu().numpy() - qo).max() < 2e-5 * scale, t
        # ---- both envs take the GPU's action; every output bit-exact
        ref = ora.step(act, obs=("i8",))
        assert np.array_equal(bits(res["reward"].cpu().numpy()), bits(ref["reward"])), t
        assert np.array_equal(res["done"].cpu().numpy(), ref["done"]), t
        assert np.array_equal(res["mask"].cpu().numpy(), ref["mask"]), t
        assert np.array_equal(res["obs"].cpu().numpy().reshape(n, 200), ref["obs_i8"].astype(np.float32)), t
        assert np.array_equal(res["ep_score"].cpu().numpy(), ref["ep_score"]), t
        # ---- store! in env order (utils.jl:244-256, 267-277)
        for i in range(n):
            rpo.store({"state": state[i], "next_state": ref["obs_i8"][i], "action_idx": act[i],
                       "reward": ref["reward"][i], "done": ref["done"][i], "mask": ref["mask"][i]})
        assert len(rb) == len(rpo) and rb.position == rpo.position, t
        # ---- target lines (utils.jl:448-451) from the oracle's Float64 t_net on the same next states
        ns = torch.from_numpy(ref["obs_i8"].reshape(n, 2, 10, 10).astype(np.float64)).to(dev)
        qn = t64(ns).cpu().numpy()
        qn_m = np.where(ref["mask"] != 0, -100.0, qn)
        y = ref["reward"].astype(np.float64) + 0.97 * qn_m.max(1) * (1 - ref["done"].astype(np.float64))
        assert np.abs(res["target"].cpu().numpy() - y).max() < 1e-4 * max(1.0, np.abs(qn).max()), t
        # the acting state of the next step: (init, init) for the envs that were re-initialised
        if t % 25 == 0 or t == steps - 1:
            assert np.array_equal(ro.state.cpu().numpy().reshape(n, 200), ora.state("i8").astype(np.float32)), t
    assert near_ties <= max(2, n * steps // 2000), near_ties
    assert env.count_errors() == 0
    # ---- the ring holds exactly what sequential store! calls produce
    L = len(rpo)
    idx = np.concatenate([np.arange(min(L, 3000)), rng.integers(0, L, 1000)]).astype(np.int64)
    got = rb.stack_exp(torch.from_numpy(idx).to(dev))
    want = rpo.stack_exp(idx)
    for k in ("states", "next_states"):
        assert np.array_equal(got[k].cpu().numpy().reshape(len(idx), 200), want[k]), k
    assert np.array_equal(got["actions"].cpu().numpy(), want["actions"])
    assert np.array_equal(bits(got["rewards"].cpu().numpy()), bits(want["rewards"]))
    assert np.array_equal(got["dones"].cpu().numpy(), want["dones"])
    assert np.array_equal(got["mask"].cpu().numpy(), want["mask"])
    # the host-array form of stack_exp returns the same bytes
    hostb = rb.stack_exp_host(torch.from_numpy(idx[:64].copy()))
    for k in ("states", "next_states", "actions", "rewards", "dones", "mask"):
        assert torch.equal(hostb[k], got[k][:64].cpu()), k


def test_patch_reset_obs_all_formats():
    """snk_patch_reset_obs == snk_state for every env after a step (the next acting state), in every observation format"""
    S = pkg()
    n = 3000
    rng = np.random.default_rng(3)
    for fmt in ("f32", "i8", "i64", "packed2"):
        env = S.SnakeGame(n, auto_reset=True)
        out = env.alloc_outputs(obs=fmt, mask=False)
        seen_done = 0
        for t in range(40):
            env.step_fused(act_idx=torch.from_numpy(rng.integers(0, 3, n).astype(np.uint8)).cuda(), out=out)
            seen_done += int(out["done"].sum())
            patched = env.patch_reset_obs(out["done"], out["obs"].clone(), fmt)
            assert torch.equal(patched, env.assemble_state(fmt)), (fmt, t)
        assert seen_done > 0
        env.close()


def test_host_getters_match_device_getters():
    S = pkg()
    n = 1234
    env = S.SnakeGame(n, auto_reset=False)
    rng = np.random.default_rng(5)
    for t in range(30):
        env.step(torch.from_numpy(rng.integers(0, 3, n).astype(np.uint8)).cuda())
    for fmt in ("f32", "i8", "i64", "packed2"):
        assert torch.equal(env.assemble_state_host(fmt), env.assemble_state(fmt).cpu()), fmt
    assert torch.equal(env.virtual_step_host(), env.virtual_step().cpu())
    assert torch.equal(env.available_actions_host(), env.available_actions().cpu())
    assert torch.equal(env.score_host(), env.score.cpu())
    assert torch.equal(env.lost_host(), env.lost.cpu())
    assert int(env.lost_host().sum()) > 0
    env.reset()                                                   # after reset! the getters show the constructor state again
    init = env.assemble_state_host("i8")
    assert torch.equal(init[:, 0], init[:, 1]) and int(env.score_host().abs().sum()) == 0 and int(env.lost_host().sum()) == 0
