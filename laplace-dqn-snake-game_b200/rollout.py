"""Batched play_episode (utils.jl:198-259) + the target lines of train! (utils.jl:448-451) for N envs.

One `step()` = for every env: epsilon_greedy on q_net(state) -> step! -> virtual_step -> store! of the Experience
into the device replay ring, then t_net(next_state) and the masked max-Q target.  Env-side work per step: ONE fused
kernel (select/step/mask/obs/store).  Its next_state output is the next step's acting state for every env that did not
finish; the rows of envs that were re-initialised are patched to (init, init) by snk_patch_reset_obs at the start of
the next step (a few % of the envs), so the 800 B/env state is never expanded a second time.
"""
import torch

from . import masked_target


class Rollout:
    def __init__(self, env, q_net, t_net=None, replay=None, epsilon=0.05):
        self.env, self.q_net, self.t_net, self.replay, self.epsilon = env, q_net, t_net or q_net, replay, float(epsilon)
        self.out = env.alloc_outputs(obs="f32", mask=True, ep_stats=True, act=True)
        env.assemble_state("f32", out=self.out["obs"])            # the acting state of the first step
        self._patch = False                                       # out["obs"] still holds terminal pairs of reset envs
        self.targets = None

    @property
    def state(self):
        """(N,2,10,10) f32: the state the next action is chosen in (assemble_state!, utils.jl:135-139)"""
        if self._patch:
            self.env.patch_reset_obs(self.out["done"], self.out["obs"])
            self._patch = False
        return self.out["obs"]

    def step(self, u=None, ridx=None):
        """Returns dict(reward, done, mask, obs (next_state: the terminal pair for an env that just lost; valid until the
        next step()), act_idx, q, target)."""
        env = self.env
        q = self.q_net(self.state)                                # (N,3) == Julia (3,N), utils.jl:165
        env.step_fused(q=q, eps=self.epsilon, u=u, ridx=ridx, out=self.out, replay=self.replay)
        self._patch = env.auto_reset
        q_next = self.t_net(self.out["obs"])                      # t_net(next_states), utils.jl:448
        self.targets = masked_target(q_next, self.out["mask"], self.out["reward"], self.out["done"])
        res = dict(self.out)
        res["target"] = self.targets
        res["q"] = q
        res["q_next"] = q_next
        return res
