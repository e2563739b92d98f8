"""Batched play_episode (utils.jl:198-259) + the target lines of train! (utils.jl:448-451) for N envs.

One `step()` = for every env: epsilon_greedy on q_net(state) -> step! -> virtual_step -> store! of the Experience
into the device replay ring, then t_net(next_state) and the masked max-Q target.  The env-side work is two
kernel launches (state view + fused select/step/mask/obs/store); the networks are whatever callable is passed
(QNet.forward_torch = library path, or the native kernel when available).
"""
import torch

from . import masked_target


class Rollout:
    def __init__(self, env, q_net, t_net=None, replay=None, epsilon=0.05):
        self.env, self.q_net, self.t_net, self.replay, self.epsilon = env, q_net, t_net or q_net, replay, float(epsilon)
        self.out = env.alloc_outputs(obs="f32", mask=True, ep_stats=True, act=True)
        self.state = env.assemble_state("f32")
        self.targets = None

    def step(self, u=None, ridx=None):
        """Returns dict(reward, done, mask, obs (next_state), act_idx, target)."""
        env = self.env
        env_state = self.state                                    # (N,2,10,10): state the action is chosen in
        q = self.q_net(env_state)                                 # (N,3) == Julia (3,N)
        env.step_fused(q=q, eps=self.epsilon, u=u, ridx=ridx, out=self.out, replay=self.replay)
        q_next = self.t_net(self.out["obs"])                      # t_net(next_states), utils.jl:448
        self.targets = masked_target(q_next, self.out["mask"], self.out["reward"], self.out["done"])
        self.state = env.assemble_state("f32")                    # (init, init) for envs that just reset
        res = dict(self.out)
        res["target"] = self.targets
        res["q"] = q
        return res
