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


# ------------------------------------------------------------------ device-stepped env.step and the graphed collector
def _twin_envs(env_cls, task, N, dev, **over):
    from hcr_genesis_lr_cl_b200 import task_spec as TS
    from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for
    spec = TS.PRESETS[task](**over)
    envs = []
    for _ in range(2):
        e = env_cls(spec, N, dev, terrain=terrain_for(spec))
        e.reset()
        envs.append(e)
    ep = envs[0].episode_length_buf
    ep[0::3] = int(envs[0].max_episode_length) - 4 - (torch.arange(len(ep[0::3]), device=ep.device) % 5).to(ep.dtype)    # resets inside the window
    envs[1].episode_length_buf = ep.clone()
    return envs


def _same_state(a, b, what):
    sa, sb = a.simulator.get_state(), b.simulator.get_state()
    for k in sa:
        if k in ("dyn_cost", "dyn_order"):
            continue
        if k == "stats":            # float atomics over the envs of a step: the order of the additions is not fixed (logging only)
            assert np.allclose(sa[k], sb[k], rtol=1e-5, atol=1e-6), f"{what}: buffer {k} differs"
            continue
        assert np.array_equal(sa[k], sb[k], equal_nan=True), f"{what}: buffer {k} differs"


def _device_stepped_equals_host_stepped(env_cls, task, N, dev, steps):
    """b200_env_step_device (per-step scalars in device memory, host draws re-drawn on the device) == b200_env_step."""
    a, b = _twin_envs(env_cls, task, N, dev)
    b.sync_step_state()
    g = torch.Generator().manual_seed(5)
    saw_reset = 0
    for t in range(steps):
        act = torch.randn(N, a.num_actions, generator=g).to(dev)
        ra, rb = a.step(act), b.step_device(act.clone())
        for i, (x, y) in enumerate(zip(ra[:-1], rb[:-1])):
            if x is not None:
                bad = (x != y).nonzero()
                assert torch.equal(x, y), (f"{task}: returned tensor {i} differs at step {t}: {bad.shape[0]} entries, first {bad[0].tolist()}, "
                                           f"{x[tuple(bad[0])].item()!r} vs {y[tuple(bad[0])].item()!r}")
        _same_state(a, b, f"{task} step {t}")
        saw_reset += int(a.reset_buf.sum())
        for k, v in a.extras["episode"].items():
            if torch.is_tensor(v):
                assert torch.allclose(v, b.extras["episode"][k], rtol=1e-5, atol=1e-7), k
    assert saw_reset > 0
    b.device_steps_done(0)
    assert a.common_step_counter == b.common_step_counter and a.simulator._hist_count == b.simulator._hist_count
    st = b.simulator.step_state.cpu().numpy()
    assert int(st[0]) == b.common_step_counter and int(st[1]) == b.simulator._hist_count


@pytest.mark.parametrize("task", ["go2_ts", "tron1_pf_ee", "go2_wtw"])
def test_device_stepped_step_equals_host_stepped_emulated(task):
    from emu_backend import EmuFusedLeggedEnv
    _device_stepped_equals_host_stepped(EmuFusedLeggedEnv, task, 8, "cpu", steps=6)


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["go2", "go2_ts", "go2_cat", "tron1_pf_ee", "go2_wtw", "go2_cts"])
def test_device_stepped_step_equals_host_stepped_gpu(task):
    from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv
    _device_stepped_equals_host_stepped(FusedLeggedEnv, task, 256, "cuda:0", steps=40)


class _Alg:
    """ppo.py:91-103 shaped stand-in (capture-safe: no host read in act)."""

    def __init__(self, no, npv, na, dev, T, N):
        from types import SimpleNamespace
        mk = lambda i, o: torch.nn.Sequential(torch.nn.Linear(i, 64), torch.nn.ELU(), torch.nn.Linear(64, o)).to(dev)
        self.actor, self.critic, self.std, self.gamma = mk(no, na), mk(npv, 1), torch.ones(na, device=dev), 0.99
        self.transition = SimpleNamespace(clear=lambda: None)
        self.actor_critic = SimpleNamespace(is_recurrent=False)
        z = lambda *s: torch.zeros(*s, device=dev)
        self.storage = SimpleNamespace(observations=z(T, N, no), privileged_observations=z(T, N, npv), rewards=z(T, N, 1), actions=z(T, N, na),
                                       dones=z(T, N, 1).byte(), actions_log_prob=z(T, N, 1), values=z(T, N, 1), mu=z(T, N, na), sigma=z(T, N, na),
                                       num_transitions_per_env=T, step=0)

    def act(self, obs, critic_obs):
        tr = self.transition
        mean = self.actor(obs)
        dist = torch.distributions.Normal(mean, mean * 0. + self.std, validate_args=False)
        tr.actions = dist.sample()
        tr.values = self.critic(critic_obs)
        tr.actions_log_prob = dist.log_prob(tr.actions).sum(dim=-1)
        tr.action_mean, tr.action_sigma = mean, dist.stddev
        return tr.actions


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["go2_ts", "go2_cat"])
def test_graphed_collector_replays_the_rollout(task):
    """GraphedRolloutCollector: T x (policy, storage copies, device-stepped env step) captured once, replayed per rollout.  The
    env side of every replayed rollout must be what a host-stepped twin env makes of the SAME actions (read back from the
    storage the graph filled): observations, rewards (with the time-out bootstrapping), dones, and the full env state."""
    from hcr_genesis_lr_cl_b200.fused_env import FusedLeggedEnv
    from hcr_genesis_lr_cl_b200.rollout import GraphedRolloutCollector
    N, T, dev = 256, 8, "cuda:0"
    a, b = _twin_envs(FusedLeggedEnv, task, N, dev)
    torch.manual_seed(1)
    alg = _Alg(b.num_obs, b.num_privileged_obs, b.num_actions, dev, T, N)
    col = GraphedRolloutCollector(b, alg, warmup_rollouts=1)
    st = alg.storage
    launches0 = None
    for r in range(5):                        # rollout 0: eager warm-up, 1: capture + replay, 2..: replays
        col.collect()
        torch.cuda.synchronize()
        assert st.step == T
        obs0 = a.obs_buf.clone()
        assert torch.equal(st.observations[0], obs0), f"rollout {r}: slot 0 is not the observation the twin stands on"
        for t in range(T):
            a.step(st.actions[t].clone())
            nxt = st.observations[t + 1] if t + 1 < T else col.carry_obs
            assert torch.equal(nxt, a.obs_buf), f"rollout {r} step {t}: observations differ"
            exp_rew = a.rew_buf + alg.gamma * st.values[t].view(-1) * a.time_out_buf
            assert torch.equal(st.rewards[t].view(-1), torch.where(a.time_out_buf, exp_rew, a.rew_buf)), f"rollout {r} step {t}: rewards differ"
            assert torch.equal(st.dones[t].view(-1).bool(), a.reset_buf)
        _same_state(a, b, f"{task} rollout {r}")
        assert a.common_step_counter == b.common_step_counter
        if r >= 2:                            # a replay issues no launch through the library's host path
            assert b.simulator._lib.b200_launch_count(b.simulator._handle) == launches0
        launches0 = b.simulator._lib.b200_launch_count(b.simulator._handle)
        st.step = 0
    assert col._graph is not None
    assert int(col.ep_stats[2]) > 0
