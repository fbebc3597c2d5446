"""Rollout-side fusion: the collection loop of rsl_rl's ``OnPolicyRunner.learn`` with the per-step copies
around ``env.step`` removed.

The reference's loop (rsl_rl/runners/on_policy_runner.py:118-139) does, every policy step,

    actions = alg.act(obs, critic_obs)                               # stores obs / critic_obs in alg.transition
    obs, privileged_obs, rewards, dones, infos = env.step(actions)
    obs, critic_obs, rewards, dones = obs.to(device), ...            # 4 .to() calls
    alg.process_env_step(rewards, dones, infos)                      # rewards.clone(), bootstrapping, add_transitions: 9 .copy_()
    cur_reward_sum += rewards; ...; rewbuffer.extend(cur_reward_sum[new_ids][:, 0].cpu().numpy().tolist())   # 2 host syncs

``FusedRolloutCollector`` keeps ``alg`` (rsl_rl's PPO, unchanged), its ``RolloutStorage`` (unchanged) and the policy in
PyTorch, and lets the env kernel write where those copies would have landed (``b200_set_rollout_targets``):

* the observation of step t goes straight into ``storage.observations[t + 1]`` / ``storage.privileged_observations[t + 1]``
  (the last one of a rollout into a carry buffer that seeds slot 0 of the next rollout),
* ``storage.rewards[t]`` (with the time-out bootstrapping ``+ gamma * values * time_outs`` of ppo.py:107-109 applied in the
  kernel from ``storage.values[t]``) and ``storage.dones[t]`` are written by the kernel,
* episode returns / lengths are accumulated on the device; the host reads three floats per logging interval.

What still goes through PyTorch per step: the policy forward (``alg.act``) and five small ``copy_`` calls for what the
policy produced (actions, values, log-prob, mean, sigma).  Nothing is copied to the host inside the loop.

Supported observation layouts: the tasks whose ``step`` returns ``obs_buf`` / ``privileged_obs_buf`` as plain per-step rows
(``go2``, and the obs / privileged pair of ``go2_ts``, ``go2_cat``, ``go2_cts``); the frame-stack tasks hand out ring views
(fused_env.py) that their own storages copy.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._cabi import B200RolloutTargets
from .fused_env import FusedLeggedEnv


class FusedRolloutCollector:
    def __init__(self, env: FusedLeggedEnv, alg, num_steps_per_env: Optional[int] = None):
        if env.stacked or env.estimator or env.spec.obs_kind == "go2_dreamwaq":
            raise ValueError("rollout fusion covers the tasks whose obs / privileged obs are per-step rows (go2, go2_ts, go2_cat, go2_cts)")
        self.env, self.alg = env, alg
        st = alg.storage
        if st is None:
            raise ValueError("alg.init_storage(...) first (on_policy_runner.py:60-64)")
        self.T = int(num_steps_per_env or st.num_transitions_per_env)
        dev = env.simulator._tdev
        N = env.num_envs
        if st.observations.shape[1:] != (N, env.num_obs) or st.observations.dtype != torch.float32 or not st.observations.is_contiguous():
            raise ValueError("storage.observations does not have the env's [T, N, num_obs] fp32 layout")
        self.has_priv = st.privileged_observations is not None
        if self.has_priv and st.privileged_observations.shape[1:] != (N, env.num_privileged_obs):
            raise ValueError("storage.privileged_observations does not match the env's privileged obs width")
        self.carry_obs = torch.zeros(N, env.num_obs, device=dev)
        self.carry_priv = torch.zeros(N, env.num_privileged_obs, device=dev) if self.has_priv else None
        self.ep_return = torch.zeros(N, device=dev)
        self.ep_length = torch.zeros(N, device=dev)
        self.ep_stats = torch.zeros(3, device=dev)
        self.bootstrap = bool(env.spec.send_timeouts)
        self._primed = False

    def prime(self) -> None:
        """Seed the carry with the env's current observation (``env.get_observations()`` in the runner, :102-105)."""
        self.carry_obs.copy_(self.env.obs_buf)
        if self.has_priv:
            self.carry_priv.copy_(self.env.privileged_obs_buf)
        self._primed = True

    @property
    def last_obs(self):
        """(obs, privileged obs) the last collected step ended on: the env's own attributes point at their alternate buffers
        during a collected rollout, the rows themselves are in the storage slabs / this carry."""
        return self.carry_obs, self.carry_priv

    def _targets(self, t: int) -> B200RolloutTargets:
        st, last = self.alg.storage, t + 1 == self.T
        obs_dst = self.carry_obs if last else st.observations[t + 1]
        priv_dst = (self.carry_priv if last else st.privileged_observations[t + 1]) if self.has_priv else None
        return B200RolloutTargets(
            obs=obs_dst.data_ptr(), privileged_obs=priv_dst.data_ptr() if priv_dst is not None else None,
            rewards=st.rewards[t].data_ptr(), dones=st.dones[t].data_ptr(),
            values=st.values[t].data_ptr() if self.bootstrap else None, gamma=float(self.alg.gamma),
            ep_return=self.ep_return.data_ptr(), ep_length=self.ep_length.data_ptr(), ep_stats=self.ep_stats.data_ptr())

    def _env_step(self, actions):
        self.env.step(actions)

    @torch.inference_mode()
    def collect(self):
        """One rollout of T policy steps into ``alg.storage``; returns the critic observation for ``alg.compute_returns``."""
        if not self._primed:
            self.prime()
        self._rollout()
        return self.carry_priv if self.has_priv else self.carry_obs

    def _rollout(self):
        env, alg, st = self.env, self.alg, self.alg.storage
        st.observations[0].copy_(self.carry_obs)                     # one copy per ROLLOUT: the observation the last one ended on
        if self.has_priv:
            st.privileged_observations[0].copy_(self.carry_priv)
        for t in range(self.T):
            obs = st.observations[t]
            critic = st.privileged_observations[t] if self.has_priv else obs
            actions = alg.act(obs, critic)                           # rsl_rl PPO.act, unchanged (ppo.py:91-103)
            tr = alg.transition
            st.actions[t].copy_(tr.actions)
            st.values[t].copy_(tr.values)
            st.actions_log_prob[t].copy_(tr.actions_log_prob.view(-1, 1))
            st.mu[t].copy_(tr.action_mean)
            st.sigma[t].copy_(tr.action_sigma)
            env.set_rollout_targets(self._targets(t))
            self._env_step(actions)                                  # obs / priv / rewards / dones land in the slabs
            st.step += 1
            tr.clear()
            if alg.actor_critic.is_recurrent:
                alg.actor_critic.reset(st.dones[t].view(-1))

    def episode_statistics(self, reset: bool = True):
        """(mean return, mean length, episodes) of the episodes that ended since the last call -- the one host read of the
        collection side, at logging time (on_policy_runner.py:201-204 reads ``statistics.mean(rewbuffer)`` there)."""
        s = self.ep_stats.tolist()
        if reset:
            self.ep_stats.zero_()
        n = max(s[2], 1.0)
        return s[0] / n, s[1] / n, int(s[2])


class GraphedRolloutCollector(FusedRolloutCollector):
    """The same rollout as ONE CUDA graph launch.

    ``OnPolicyRunner.learn``'s collection loop is launch-bound in PyTorch eager mode: per policy step the actor / critic MLPs,
    the Gaussian sampling and the storage copies are ~60 small kernels the host issues one by one, next to which the env step
    (three launches, ~0.2 ms on the device) waits.  The env's device-stepped entry point (``FusedLeggedEnv.step_device`` ->
    ``b200_env_step_device``) takes no per-step value from the host -- counters, the Philox step key, the frame-stack
    position, the host-drawn gait / sit-pose scalars and the curriculum ranges live in device memory -- so the whole loop of
    ``T`` x (``alg.act`` -> storage copies -> env step) is captured once with ``torch.cuda.graph`` and replayed per rollout.
    Nothing but kernels is left in the replay: results are those of ``FusedRolloutCollector`` driven through ``step_device``
    (bit-identical env side; the policy's sampling uses torch's graph-safe Philox offsets).

    Restrictions: non-recurrent policies; an even ``T`` (the env's per-step rows and the dynamics kernel's scheduling buffers
    alternate between two halves); ``alg.act`` must be capture-safe -- no host reads.  rsl_rl's own ``ActorCritic.act`` has
    two: ``torch.distributions`` argument validation (switched off for the capture, as rsl_rl intends with its
    ``Normal.set_default_validate_args``) and ``torch.normal(Tensor, Tensor)``'s ``std >= 0`` check behind ``Normal.sample``
    (replaced for the capture by the arithmetic it performs, ``randn * std + mean`` on the same generator).
    The command / behaviour curricula are applied at rollout boundaries (``FusedLeggedEnv.device_steps_done``).
    """

    def __init__(self, env: FusedLeggedEnv, alg, num_steps_per_env: Optional[int] = None, warmup_rollouts: int = 1):
        super().__init__(env, alg, num_steps_per_env)
        if self.T % 2:
            raise ValueError("graphed collection needs an even number of steps per rollout")
        if getattr(alg.actor_critic, "is_recurrent", False):
            raise ValueError("graphed collection covers non-recurrent policies")
        self._graph = None
        self._warmup_left = int(warmup_rollouts)
        self._synced = False

    def _env_step(self, actions):
        self.env.step_device(actions)

    def _sync(self):
        if not self._synced:
            self.env.sync_step_state()
            self._synced = True

    @torch.inference_mode()
    def collect(self):
        env, st = self.env, self.alg.storage
        if not self._primed:
            self.prime()
        self._sync()
        if self._warmup_left > 0:
            # eager, device-stepped: lets cuBLAS / the caching allocator / the library's side stream set themselves up outside a capture
            self._warmup_left -= 1
            self._rollout()
            env.device_steps_done(0)
        elif self._graph is None:
            from torch.distributions import Distribution, Normal
            prev, prev_sample = Distribution._validate_args, Normal.sample
            Distribution.set_default_validate_args(False)

            def sample(dist, sample_shape=torch.Size()):
                # torch.normal(Tensor, Tensor) checks std >= 0 with a host read (not capturable); its arithmetic is
                # randn * std + mean on the same generator, which is what gets captured instead
                shape = dist._extended_shape(sample_shape)
                with torch.no_grad():
                    return torch.randn(shape, dtype=dist.loc.dtype, device=dist.loc.device).mul_(dist.scale.expand(shape)).add_(dist.loc.expand(shape))

            Normal.sample = sample
            try:
                g = torch.cuda.CUDAGraph()
                counters = (env.common_step_counter, env.simulator._hist_count, env._pp_i, st.step)
                torch.cuda.synchronize(env.simulator._tdev)
                with torch.cuda.graph(g):
                    self._rollout()
                # the capture ran the python side of T steps but no kernel: rewind the host's view and replay for real
                env.common_step_counter, env.simulator._hist_count, env._pp_i, st.step = counters
                self._graph = g
            finally:
                Distribution.set_default_validate_args(prev)
                Normal.sample = prev_sample
            self._replay()
        else:
            self._replay()
        return self.carry_priv if self.has_priv else self.carry_obs

    def _replay(self):
        self._graph.replay()
        self.alg.storage.step += self.T
        self.env.device_steps_done(self.T)
