#!/usr/bin/env python
"""Collection throughput with the policy in the loop (SURVEY 8d "full runner.learn() collection", 8f rank 2).

Two loops over the same FusedLeggedEnv (`go2_ts`, 4096 envs) and the same rsl_rl-shaped actor / critic MLPs
(45 -> 512 -> 256 -> 128 -> 12 and 99 -> 512 -> 256 -> 128 -> 1, PyTorch):

  runner     the reference's collection loop as written (on_policy_runner.py:118-139): alg.act -> env.step -> .to(device) x4 ->
             process_env_step (rewards.clone, bootstrapping, 9 copy_ into the storage) -> per-step episode book-keeping with its
             two .cpu().numpy().tolist() host reads;
  fused      rollout.FusedRolloutCollector: the env kernel writes obs / privileged obs / rewards / dones into the storage slabs,
             book-keeping stays on the device, no host read inside the loop;
  graphed    rollout.GraphedRolloutCollector: the fused loop with the device-stepped env step (b200_env_step_device), captured
             once into a CUDA graph and replayed: one launch per rollout.

The PPO / storage classes are small stand-ins with rsl_rl's attribute names (the reference tree does not travel to the GPU
box for bench purposes); tests/test_rollout_fusion.py checks the collector against the reference's own classes.

    python tools/bench_collection.py [--envs 4096] [--iters 20]
"""
import argparse
import json
import os
import sys
import time
from collections import deque
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hcr_genesis_lr_cl_b200 import build, task_spec as T  # noqa: E402
from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv  # noqa: E402
from hcr_genesis_lr_cl_b200.rollout import FusedRolloutCollector, GraphedRolloutCollector  # noqa: E402
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for  # noqa: E402


def mlp(i, o):
    return torch.nn.Sequential(torch.nn.Linear(i, 512), torch.nn.ELU(), torch.nn.Linear(512, 256), torch.nn.ELU(),
                               torch.nn.Linear(256, 128), torch.nn.ELU(), torch.nn.Linear(128, o))


class Storage:
    """rollout_storage.py:46-103 (the fields the collection touches)."""

    def __init__(self, T_, N, no, npv, na, dev):
        z = lambda *s: torch.zeros(*s, device=dev)
        self.observations, self.privileged_observations = z(T_, N, no), z(T_, N, npv)
        self.rewards, self.actions, self.dones = z(T_, N, 1), z(T_, N, na), z(T_, N, 1).byte()
        self.actions_log_prob, self.values, self.mu, self.sigma = z(T_, N, 1), z(T_, N, 1), z(T_, N, na), z(T_, N, na)
        self.num_transitions_per_env, self.step = T_, 0

    def add_transitions(self, tr):
        s = self.step
        self.observations[s].copy_(tr.observations); self.privileged_observations[s].copy_(tr.critic_observations)
        self.actions[s].copy_(tr.actions); self.rewards[s].copy_(tr.rewards.view(-1, 1)); self.dones[s].copy_(tr.dones.view(-1, 1))
        self.values[s].copy_(tr.values); self.actions_log_prob[s].copy_(tr.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(tr.action_mean); self.sigma[s].copy_(tr.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0


class Alg:
    """ppo.py:91-116 (act / process_env_step)."""

    def __init__(self, actor, critic, na, dev, gamma=0.99):
        self.actor, self.critic, self.std, self.gamma = actor, critic, torch.ones(na, device=dev), gamma
        self.transition = SimpleNamespace(clear=lambda: None)
        self.actor_critic = SimpleNamespace(is_recurrent=False, reset=lambda d: None)
        self.storage = None

    def act(self, obs, critic_obs):
        tr = self.transition
        mean = self.actor(obs)
        dist = torch.distributions.Normal(mean, mean * 0. + self.std, validate_args=False)
        tr.actions = dist.sample()
        tr.values = self.critic(critic_obs)
        tr.actions_log_prob = dist.log_prob(tr.actions).sum(dim=-1)
        tr.action_mean, tr.action_sigma = mean, dist.stddev
        tr.observations, tr.critic_observations = obs, critic_obs
        return tr.actions

    def process_env_step(self, rewards, dones, infos):
        tr = self.transition
        tr.rewards = rewards.clone()
        tr.dones = dones
        if "time_outs" in infos:
            tr.rewards += self.gamma * torch.squeeze(tr.values * infos["time_outs"].unsqueeze(1), 1)
        self.storage.add_transitions(tr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--horizon", type=int, default=24)
    ap.add_argument("--task", default="go2_ts")
    args = ap.parse_args()
    dev = "cuda:0"
    build.build()
    spec = T.PRESETS[args.task]()
    N, H = args.envs, args.horizon
    out = {}
    for mode in ("runner", "fused", "graphed"):
        env = FusedLeggedEnv(spec, N, dev, terrain=terrain_for(spec))
        env.reset()
        torch.manual_seed(0)
        npv = env.num_privileged_obs or env.num_obs
        alg = Alg(mlp(env.num_obs, spec.num_actions).to(dev), mlp(npv, 1).to(dev), spec.num_actions, dev)
        alg.storage = Storage(H, N, env.num_obs, npv, spec.num_actions, dev)
        col = FusedRolloutCollector(env, alg) if mode == "fused" else (GraphedRolloutCollector(env, alg) if mode == "graphed" else None)
        rewbuffer, lenbuffer = deque(maxlen=100), deque(maxlen=100)
        cur_reward_sum, cur_episode_length = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
        for _ in range(300 // H + 1):                       # steady-state workload first
            with torch.inference_mode():
                for _ in range(H):
                    env.step(torch.randn(N, spec.num_actions, device=dev))
        times = []
        for it in range(args.iters + 3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if col is not None:
                col.collect()
                col.episode_statistics()
            else:
                obs, critic_obs = env.obs_buf, env.privileged_obs_buf
                with torch.inference_mode():
                    for i in range(H):
                        actions = alg.act(obs, critic_obs)
                        o = env.step(actions)
                        obs, critic_obs, rewards, dones, infos = o[0], o[1], o[-3], o[-2], o[-1]
                        obs, critic_obs, rewards, dones = obs.to(dev), critic_obs.to(dev), rewards.to(dev), dones.to(dev)
                        alg.process_env_step(rewards, dones, infos)
                        cur_reward_sum += rewards
                        cur_episode_length += 1
                        new_ids = (dones > 0).nonzero(as_tuple=False)
                        rewbuffer.extend(cur_reward_sum[new_ids][:, 0].cpu().numpy().tolist())
                        lenbuffer.extend(cur_episode_length[new_ids][:, 0].cpu().numpy().tolist())
                        cur_reward_sum[new_ids] = 0
                        cur_episode_length[new_ids] = 0
            alg.storage.clear()
            torch.cuda.synchronize()
            if it >= 3:                      # graphed: warm-up rollout, capture, first replay
                times.append(time.perf_counter() - t0)
        times.sort()
        med = times[len(times) // 2]
        out[mode] = {"ms_per_policy_step": 1e3 * med / H, "env_substeps_per_s": N * spec.decimation * H / med,
                     "host_reads_per_policy_step": 2 if mode == "runner" else 0,
                     "launches_per_rollout": 1 if mode == "graphed" else None}
        del env
    out["speedup"] = out["runner"]["ms_per_policy_step"] / out["fused"]["ms_per_policy_step"]
    out["speedup_graphed"] = out["runner"]["ms_per_policy_step"] / out["graphed"]["ms_per_policy_step"]
    print(json.dumps({"task": args.task, "envs": N, "horizon": H, "policy": "MLP 512-256-128 actor + critic (PyTorch)", **out}))


if __name__ == "__main__":
    main()
