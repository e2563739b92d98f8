#!/usr/bin/env python3
"""The same data path for a trainer whose Q-net stays on the HOST (the reference as it is: Flux on the CPU): per step the
Q-values of every game go up, one 24-byte record per game comes down (SNK_OBS_BITS: the next state as two bit-boards, reward,
lost, next_is_suicidal, the action taken), every Experience is store!d into the device replay ring, and the minibatch is
gathered as Float32 where the reference casts (stack_exp, utils.jl:343-383).  Needs a B200.

  q = q_net(game.state)                   -> host array (N, 3); here a stand-in function of the decoded state
  epsilon_greedy + step! + virtual_step   -> SnakeGame.step_fused_host(host, q=True, eps, replay)
  game.state / reward / lost / mask       -> unpack_bits(host["obs"])
  sample(rpb) + stack_exp                 -> ReplayBuffer.stack_exp_host(idx)
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    S = graft.load_package()
    n = args.envs
    env = S.SnakeGame(n, auto_reset=True)
    rb = S.ReplayBuffer(capacity=50000, batch_size=64)
    host = {"obs_fmt": "bits", "q": S.pinned_empty((n, 3), torch.float32), "obs": S.pinned_empty((n, 24), torch.uint8)}
    host["q"].zero_()
    episodes, ret, t_step = 0, 0.0, 0.0
    for t in range(args.steps):
        t0 = time.perf_counter()
        env.step_fused_host(host, q=True, eps=0.05, replay=rb)        # asynchronous: returns once everything is enqueued
        env.sync()                                                    # the records are in host memory
        t_step += time.perf_counter() - t0
        d = S.unpack_bits(host["obs"])                                # state (N,2,10,10) int8, reward, done, mask, action
        episodes += int(d["done"].sum())
        ret += float(d["reward"].sum())
        # stand-in for the host Q-net: prefer moves that are not suicidal (a real trainer evaluates Flux on d["state"])
        host["q"].copy_(1.0 - d["mask"].float() + 0.01 * torch.rand(n, 3))
    idx = torch.randint(0, len(rb), (64,), dtype=torch.int64)
    batch = rb.stack_exp_host(idx)                                    # Float32 (64,2,10,10) states / next_states, actions, ...
    print("env-steps %d in %.1f ms of step+copy time (%.3g env-steps/s), episodes finished %d, mean reward/step %.4f, replay %d/%d, "
          "minibatch states %s, env errors %d" % (n * args.steps, 1e3 * t_step, n * args.steps / t_step, episodes,
                                                   ret / (n * args.steps), len(rb), rb.capacity, tuple(batch["states"].shape),
                                                   env.count_errors()))


if __name__ == "__main__":
    main()
