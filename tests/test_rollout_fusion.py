"""Rollout-side fusion: `FusedRolloutCollector` fills rsl_rl's own `RolloutStorage` (imported from the
reference, unchanged) through the env kernel's extra destinations; the slabs must equal what the reference's collection
loop -- `alg.act` -> `env.step` -> `alg.process_env_step` / `add_transitions` (on_policy_runner.py:118-139,
ppo.py:91-116, rollout_storage.py:89-103) -- leaves behind on the same env, seed and policy."""
import os
import sys

import numpy as np
import pytest
import torch

from plugin_util import reference_root

needs_reference = pytest.mark.skipif(reference_root() is None, reason="no reference tree (/root/reference or baseline/_ref)")


def _rsl_rl():
    root = reference_root()
    if root not in sys.path:
        sys.path.insert(0, root)
    from rsl_rl.algorithms import PPO
    from rsl_rl.modules import ActorCritic
    return PPO, ActorCritic


def _make(env_cls, task, N, dev, T):
    from hcr_genesis_lr_cl_b200 import task_spec as TS
    from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for
    PPO, ActorCritic = _rsl_rl()
    spec = TS.PRESETS[task]()
    env = env_cls(spec, N, dev, terrain=terrain_for(spec))
    env.reset()
    torch.manual_seed(3)
    n_priv = env.num_privileged_obs
    ac = ActorCritic(env.num_obs, n_priv if n_priv is not None else env.num_obs, env.num_actions,
                     actor_hidden_dims=[64, 32], critic_hidden_dims=[64, 32], activation="elu", init_noise_std=1.0)
    alg = PPO(ac, device=dev)
    alg.init_storage(N, T, [env.num_obs], [n_priv], [env.num_actions])
    return env, alg


def _obs_pair(env, out=None):
    return env.obs_buf, (env.privileged_obs_buf if env.num_privileged_obs is not None else None)


def _compare(env_cls, task, N, dev, T=24, rollouts=2):
    from hcr_genesis_lr_cl_b200.rollout import FusedRolloutCollector
    env_a, alg_a = _make(env_cls, task, N, dev, T)
    env_b, alg_b = _make(env_cls, task, N, dev, T)
    ep = env_a.episode_length_buf
    ep[0::3] = int(env_a.max_episode_length) - 5 - (torch.arange(len(ep[0::3]), device=ep.device) % 9).to(ep.dtype)     # time-outs inside the rollout
    env_b.episode_length_buf = ep.clone()
    col = FusedRolloutCollector(env_b, alg_b)
    cur_ret, cur_len, ret_sum, len_sum, n_done = torch.zeros(N, device=dev), torch.zeros(N, device=dev), 0.0, 0.0, 0
    for r in range(rollouts):
        # ---- A: the reference's loop
        torch.manual_seed(100 + r)
        obs, priv = _obs_pair(env_a)
        with torch.inference_mode():
            for t in range(T):
                critic = priv if priv is not None else obs
                actions = alg_a.act(obs, critic)
                out = env_a.step(actions)
                rew, dones, infos = out[-3], out[-2], out[-1]
                obs, priv = _obs_pair(env_a)
                alg_a.process_env_step(rew, dones, infos)
                cur_ret += rew; cur_len += 1
                ids = (dones > 0).nonzero(as_tuple=False)
                ret_sum += float(cur_ret[ids].sum()); len_sum += float(cur_len[ids].sum()); n_done += int(ids.numel())
                cur_ret[ids] = 0; cur_len[ids] = 0
        # ---- B: the fused collector
        torch.manual_seed(100 + r)
        last_critic = col.collect()
        sa, sb = alg_a.storage, alg_b.storage
        for name in ("observations", "privileged_observations", "actions", "rewards", "dones", "values", "actions_log_prob", "mu", "sigma"):
            xa, xb = getattr(sa, name), getattr(sb, name)
            if xa is None:
                assert xb is None
                continue
            assert torch.equal(xa, xb), f"rollout {r}: storage.{name} differs (max {float((xa.float() - xb.float()).abs().max()):.3e})"
        assert int(sa.dones.sum()) > 0 and sa.step == sb.step == T
        ref_last = env_a.privileged_obs_buf if env_a.num_privileged_obs is not None else env_a.obs_buf
        assert torch.equal(last_critic, ref_last)
        sa.clear(); sb.clear()
    m_ret, m_len, n = col.episode_statistics()
    assert n == n_done and n > 0
    assert abs(m_ret - ret_sum / n_done) <= 1e-4 * abs(ret_sum / n_done) + 1e-5 and abs(m_len - len_sum / n_done) < 1e-3


@needs_reference
@pytest.mark.parametrize("task", ["go2", "go2_ts"])
def test_fused_collector_fills_the_reference_storage_emulated(task):
    from emu_backend import EmuFusedLeggedEnv
    _compare(EmuFusedLeggedEnv, task, 12, "cpu", T=8, rollouts=2)


@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize("task", ["go2", "go2_ts", "go2_cat"])
def test_fused_collector_fills_the_reference_storage_gpu(task):
    from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv
    _compare(FusedLeggedEnv, task, 256, "cuda:0", T=24, rollouts=2)
