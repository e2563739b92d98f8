#!/usr/bin/env python3
"""What a user of the reference does with this library: the data path of train! (utils.jl:420-494) without the
learner.  Needs a B200.

  Trainer(...)                      -> SnakeGame(N), ReplayBuffer(50000), QNet (from a BSON checkpoint or Flux's default init)
  fill_buffer! / play_episode       -> Rollout.step(): eps-greedy on q_net, step!, virtual_step, store!  (one kernel + the net)
  sample + stack_exp                -> ReplayBuffer.sample()
  q_next[mask] .= -100; max; target -> masked_target(t_net(next_states), mask, rewards, dones)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--bson", default=None, help="trainers/<name>.bson of a two-frame Q-net (optional)")
    ap.add_argument("--epsilon", type=float, default=0.05)
    args = ap.parse_args()
    S = graft.load_package()
    dev = torch.device("cuda", 0)
    if args.bson:
        q_layers, t_layers = S.bson_io.load_trainer_nets(args.bson)
    else:
        q_layers = t_layers = S.qnet.glorot_layers(seed=0)           # Flux default init (structs.jl:127-139)
    q_net = S.qnet.QNet(q_layers, dev, precision="f32")
    t_net = S.qnet.QNet(t_layers, dev, precision="f32")
    env = S.SnakeGame(args.envs, auto_reset=True)
    rb = S.ReplayBuffer(capacity=50000, batch_size=64)
    ro = S.rollout.Rollout(env, q_net, t_net, rb, epsilon=args.epsilon)
    done_total, reward_total = 0, 0.0
    for _ in range(args.steps):
        res = ro.step()
        done_total += int(res["done"].sum())
        reward_total += float(res["reward"].sum())
    batch = rb.sample()                                               # sample(rpb) + stack_exp
    y = S.masked_target(t_net(batch["next_states"]), batch["mask"], batch["rewards"], batch["dones"])
    q_sel = q_net(batch["states"]).gather(1, batch["actions"].long()[:, None])[:, 0]
    huber = torch.nn.functional.huber_loss(q_sel.double(), y)         # Flux.huber_loss(q_pred_selected, q_target), utils.jl:456
    print("env-steps %d, episodes finished %d, mean reward/step %.4f, replay %d/%d, minibatch Huber loss %.4f, env errors %d"
          % (args.envs * args.steps, done_total, reward_total / (args.envs * args.steps), len(rb), rb.capacity,
             float(huber), env.count_errors()))


if __name__ == "__main__":
    main()
