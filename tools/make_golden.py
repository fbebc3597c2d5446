#!/usr/bin/env python
"""Record golden vectors from the reference's own environment code.

Runs LeggedGym-Ex's unmodified Python classes (LeggedRobot / GO2 / Go2TS + GenesisSimulator, imported
from /root/reference) over the CPU physics oracle behind a stand-in `genesis` module
(oracle/stubs/genesis), with every random draw replaced by the counter-based Philox stream the fused
kernel uses (oracle/philox.py).  Writes tests/golden/<name>.npz holding

  * the full environment state before the recorded window (``s0/*``),
  * per step: the actions, the state right after the decimated physics loop (``phys/*`` -- what the
    engine hands to post_physics_step) and everything the reference computed from it (``out/*``).

Build-container tooling (needs /root/reference); the fixtures it writes travel to the GPU box.

Usage: python tools/make_golden.py [--task go2_ts] [--envs 32] [--steps 10] [--name go2_ts_n32]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hcr_genesis_lr_cl_b200 import task_spec as T  # noqa: E402
from oracle import philox  # noqa: E402,F401
from oracle.inject import Injector  # noqa: E402
from oracle.ref_harness import make_env  # noqa: E402


def pack_gait(env):
    """theta[4], gait_time, phi, gait_period, base_height_t, foot_clearance_t, pitch_t, clock[8] (B200Buffers.gait_state)."""
    n = lambda t: t.detach().cpu().numpy().astype(np.float32)
    N = env.num_envs
    g = np.zeros((N, 20), np.float32)
    th, clk = n(env.theta), n(env.clock_input)
    g[:, 0:th.shape[1]] = th
    g[:, 4:5], g[:, 5:6], g[:, 6:7] = n(env.gait_time), n(env.phi), n(env.gait_period)
    for k, name in ((7, "base_height_target"), (8, "foot_clearance_target"), (9, "pitch_target")):
        if hasattr(env, name):
            g[:, k:k + 1] = n(getattr(env, name))
    g[:, 10:10 + clk.shape[1]] = clk
    return g


def snapshot(env, spec, sum_names):
    """Reference env -> EnvOracle/B200 state dictionary (fp32, dof order = cfg.asset.dof_names)."""
    sim, rob = env.simulator, env.simulator._robot
    didx = [d - 6 for d in sim._dof_indices]
    n = lambda t: t.detach().cpu().numpy().copy()
    st = dict(
        base_pos=rob.state[:, 0:3], base_quat_wxyz=rob.state[:, 3:7], base_lin_w=rob.state[:, 7:10],
        base_ang_w=rob.state[:, 10:13], q=rob.q[:, didx], qd=rob.qd[:, didx],
        friction=n(sim._friction_values), added_mass=n(sim._added_base_mass), com_bias=n(sim._base_com_bias),
        kp_scale=n(sim._kp_scale), kd_scale=n(sim._kd_scale), joint_armature=n(sim._joint_armature),
        joint_friction=n(sim._joint_friction), joint_damping=n(sim._joint_damping), rand_push_vels=n(sim._rand_push_vels),
        actions=n(env.actions), last_actions=n(env.last_actions), llast_actions=n(env.llast_actions), commands=n(env.commands),
        episode_length=n(env.episode_length_buf), fail_buf=n(env.fail_buf), feet_air_time=n(env.feet_air_time),
        last_contacts=n(env.last_contacts).astype(np.uint8),
        episode_sums=np.stack([n(env.episode_sums[k]) for k in sum_names], axis=1),
        env_origins=n(sim._env_origins),
        base_lin_vel=n(sim._base_lin_vel), base_ang_vel=n(sim._base_ang_vel), feet_vel=n(sim._feet_vel),
        last_dof_vel=n(sim._last_dof_vel), last_feet_vel=n(sim._last_feet_vel),
        last_base_lin_vel=n(sim._last_base_lin_vel), last_base_ang_vel=n(sim._last_base_ang_vel),
    )
    if spec.heightfield:
        st["terrain_levels"], st["terrain_types"] = n(sim._terrain_levels), n(sim._terrain_types)
    else:
        st["terrain_levels"] = st["terrain_types"] = np.zeros(env.num_envs, np.int32)
    if hasattr(env, "gait_time"):
        st["gait_state"] = pack_gait(env)
    if getattr(env.cfg.domain_rand, "randomize_ctrl_delay", False):
        st["action_queue"], st["action_delay"] = n(env.action_queue), n(env.action_delay)
    if hasattr(env, "critic_history"):                       # go2_wtw names its deques obs_history / critic_history
        st["obs_hist"] = np.concatenate([n(x) for x in env.obs_history], axis=1)
        st["critic_hist"] = np.concatenate([n(x) for x in env.critic_history], axis=1)
    if hasattr(env, "obs_history_deque"):
        st["obs_hist"] = np.concatenate([n(x) for x in env.obs_history_deque], axis=1)
        st["critic_hist"] = np.concatenate([n(x) for x in env.critic_obs_deque], axis=1)
    out = {}
    for k, v in st.items():
        v = np.asarray(v)
        out[k] = v.astype(np.int32) if v.dtype.kind in "iub" and k != "last_contacts" else (v if k == "last_contacts" else v.astype(np.float32))
    out["common_step_counter"] = np.int64(env.common_step_counter)
    if hasattr(env, "gait_period_range"):                     # go2_wtw host-side behaviour state
        out["beh_ranges"] = np.asarray([env.gait_period_range, env.base_height_target_range, env.foot_clearance_target_range,
                                        env.pitch_target_range], np.float64)
        out["num_gaits"] = np.int64(env.num_gaits)
    out["cmd_range_x"] = np.asarray(env.command_ranges["lin_vel_x"], np.float64)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="go2_ts")
    ap.add_argument("--envs", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--name", default=None)
    ap.add_argument("--contact-links", nargs="*", default=None, help="override asset.contact_state_link_names (SURVEY R1)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--ctrl-delay", type=int, default=0, help="enable domain_rand.randomize_ctrl_delay with step range [0, N]")
    ap.add_argument("--gait-smooth", action="store_true", help="periodic_reward_framework.gait_function_type = 'smooth' (von Mises CDF indicator)")
    ap.add_argument("--all-rewards", action="store_true", help="give every _reward_* term the task class defines a non-zero scale")
    ap.add_argument("--oob", action="store_true", help="throw a few robots off the terrain inside the window (OOB teleport, genesis_simulator.py:612-628)")
    args = ap.parse_args()
    import torch

    def edit(cfg):
        if args.contact_links is not None:
            cfg.asset.contact_state_link_names = list(args.contact_links)
        if args.ctrl_delay > 0:
            cfg.domain_rand.randomize_ctrl_delay = True
            cfg.domain_rand.ctrl_delay_step_range = [0, args.ctrl_delay]
        if args.gait_smooth:
            cfg.rewards.periodic_reward_framework.gait_function_type = "smooth"
        if args.all_rewards:
            # every term this task class implements and the fused kernel knows, with a non-zero scale (the shipped configs
            # never enable dof_pos_stand_still, dof_vel_stand_still, termination, thigh_pos, ...); small magnitudes so that the
            # only_positive_rewards clip does not flatten everything
            import legged_gym.envs  # noqa: F401
            from legged_gym.utils.task_registry import task_registry as _reg
            klass = _reg.get_task_class(args.task)
            if not hasattr(cfg.rewards, "about_landing_threshold"):      # read by _reward_foot_landing_vel, missing from the rough-terrain configs
                cfg.rewards.about_landing_threshold = 0.03
            for k, name in enumerate(T.REWARD_TERMS):
                if hasattr(klass, "_reward_" + name) and name not in ("torque_limits",) and float(getattr(cfg.rewards.scales, name, 0.0) or 0.0) == 0.0:
                    setattr(cfg.rewards.scales, name, (-1.0 if name != "keep_balance" else 1.0) * 1e-3 * (1 + k % 5))

    env, cfg, _ = make_env(args.task, args.envs, cfg_edit=edit)
    spec = T.TaskSpec.from_reference_cfg(cfg, args.task)
    spec.seed = args.seed
    sim = env.simulator
    N, A = env.num_envs, env.num_actions
    g = torch.Generator().manual_seed(1234)
    env.reset()
    for _ in range(args.warmup):
        env.step(0.8 * torch.randn(N, A, generator=g))
    # spread the interesting events over the recorded window
    L = env.max_episode_length
    ep = env.episode_length_buf
    ep[:] = torch.randint(10, 400, (N,), generator=g).to(ep.dtype)
    ep[0::8] = int(L) - 3 - (torch.arange(len(ep[0::8])) % 3).to(ep.dtype)         # time-outs
    ep[1::8] = 500 - 2 - (torch.arange(len(ep[1::8])) % 4).to(ep.dtype)            # command resampling
    env.fail_buf[2::8] = 3                                                         # one more fail -> reset
    rob = sim._robot
    flip = np.arange(2, N, 8)
    rob.state[flip, 3:7] = np.array([0.0, 1.0, 0.0, 0.0])                          # upside down
    rob.state[flip, 2] += 0.3 if A == 12 else 0.8
    env.common_step_counter = spec.push_interval * 3 - args.steps // 2             # a push inside the window
    # a few envs with a (near-)zero command: stand-still rewards / CaT stand-still constraint
    from legged_gym.utils.math_utils import quat_apply
    q_xyzw = torch.from_numpy(np.concatenate([rob.state[:, 4:7], rob.state[:, 3:4]], axis=1)).float()
    fwd = quat_apply(q_xyzw, torch.tensor([[1.0, 0.0, 0.0]]).repeat(N, 1))
    env.commands[3::8, :3] = 0.0
    env.commands[3::8, 3] = torch.atan2(fwd[3::8, 1], fwd[3::8, 0])
    if hasattr(env, "num_gaits"):                                                  # exercise every gait and real ranges
        env.num_gaits = env.num_gait_max
        env.gait_period_range, env.base_height_target_range = [0.35, 0.55], [0.24, 0.32]
        env.foot_clearance_target_range, env.pitch_target_range = [0.05, 0.10], [-0.2, 0.2]
        ep[4::8] = 250 - 2 - (torch.arange(len(ep[4::8])) % 4).to(ep.dtype)        # behaviour resampling inside the window
    env.step(torch.zeros(N, A))                                                    # make API buffers consistent
    # the reference keeps episode_sums as a dict (alphabetical keys, CaT adds its cstr_* keys lazily); the recording stacks the
    # columns in the descriptor's order, which differs only in that "termination" comes after the other reward terms
    sum_names = spec.episode_sum_names()
    assert sorted(sum_names) == sorted(env.episode_sums.keys()), (sum_names, list(env.episode_sums.keys()))
    inj = Injector(env, spec.seed)
    inj.install()

    s0 = snapshot(env, spec, sum_names)
    rec = {f"s0/{k}": v for k, v in s0.items()}
    feet = sim._feet_indices
    didx = [d - 6 for d in sim._dof_indices]
    orig_step = sim.step
    phys = {}

    def sim_step(actions):
        orig_step(actions)
        lp, lv = rob.link_pos, rob.link_vel
        phys.update(base_pos=rob.state[:, 0:3].copy(), base_quat_wxyz=rob.state[:, 3:7].copy(),
                    base_lin_w=rob.state[:, 7:10].copy(), base_ang_w=rob.state[:, 10:13].copy(),
                    q=rob.q[:, didx].copy(), qd=rob.qd[:, didx].copy(), torques=sim._torques.numpy().copy(),
                    link_force=rob.link_force.copy(), feet_pos=lp[:, feet].copy(), feet_vel=lv[:, feet].copy())
    sim.step = sim_step
    n = lambda t: t.detach().cpu().numpy().copy()
    per = {}

    def put(t, k, v):
        per.setdefault(k, []).append(np.asarray(v))
    for t in range(args.steps):
        a = 0.8 * torch.randn(N, A, generator=g)
        if t == 1:
            a[0] = 150.0                                                            # exercise clip_actions
        if args.oob and t == 2:
            rob.state[5::8, 0] += 400.0                                             # off the heightfield: teleported home in post_physics_step
            rob.state[6::8, 1] -= 400.0
        ret = env.step(a.clone())
        put(t, "actions", n(a))
        for k, v in phys.items():
            put(t, "phys/" + k, v.astype(np.float32))
        o = dict(obs_buf=n(ret[0]), rew_buf=n(env.rew_buf), reset_buf=n(env.reset_buf).astype(np.uint8),
                 time_out_buf=n(env.time_out_buf).astype(np.uint8), commands=n(env.commands),
                 episode_length=n(env.episode_length_buf), fail_buf=n(env.fail_buf), feet_air_time=n(env.feet_air_time),
                 last_contacts=n(env.last_contacts).astype(np.uint8),
                 episode_sums=np.stack([n(env.episode_sums[k]) for k in sum_names], axis=1),
                 env_origins=n(sim._env_origins), dof_pos=n(sim._dof_pos), dof_vel=n(sim._dof_vel), base_pos=n(sim._base_pos),
                 base_lin_vel=n(sim._base_lin_vel), base_ang_vel=n(sim._base_ang_vel),
                 projected_gravity=n(sim._projected_gravity), base_euler=n(sim._base_euler),
                 friction=n(sim._friction_values), added_mass=n(sim._added_base_mass), com_bias=n(sim._base_com_bias),
                 kp_scale=n(sim._kp_scale), kd_scale=n(sim._kd_scale), rand_push_vels=n(sim._rand_push_vels),
                 actions_buf=n(env.actions), last_actions=n(env.last_actions), llast_actions=n(env.llast_actions),
                 measured_heights=n(sim._measured_heights), link_contact_forces=n(sim._link_contact_forces),
                 feet_pos=n(sim._feet_pos), feet_vel=n(sim._feet_vel),
                 end_q=rob.q[:, didx].astype(np.float32), end_qd=rob.qd[:, didx].astype(np.float32),
                 end_state=rob.state.astype(np.float32))
        if spec.heightfield:
            o.update(terrain_levels=n(sim._terrain_levels), height_around_feet=n(sim._height_around_feet),
                     normal_vector_around_feet=n(sim._normal_vector_around_feet))
        if spec.obtain_link_contact_states:
            o["link_contact_states"] = n(sim._link_contact_states)
        if hasattr(env, "cstr_prob"):
            o["cstr_prob"] = n(env.cstr_prob)
        if args.ctrl_delay > 0:
            o["action_queue"], o["action_delay"] = n(env.action_queue).reshape(N, -1), n(env.action_delay).astype(np.int32)
        if args.task == "go2_wtw":
            o["privileged_obs_buf"] = n(ret[1])
            o["gait_state"] = pack_gait(env)
        elif args.task == "go2_dreamwaq":                 # obs, critic stack, history, explicit labels, next state
            o["privileged_obs_buf"] = n(ret[1])
            o["explicit_labels_buf"] = n(ret[3])
            o["next_state_buf"] = n(ret[4])
        elif args.task in ("tron1_pf_ee", "go2_ee"):
            o["estimator_labels_buf"] = n(ret[1])
            o["privileged_obs_buf"] = n(ret[2])
            if hasattr(env, "gait_time"):
                o["gait_state"] = pack_gait(env)
        elif ret[1] is not None:
            o["privileged_obs_buf"] = n(ret[1])
        if hasattr(env, "obs_history") and args.task not in ("tron1_pf_ee", "go2_wtw", "go2_ee") and t in (args.steps // 2, args.steps - 1):
            rec[f"hist{t}/obs_history"] = n(env.obs_history)
            rec[f"hist{t}/critic_obs_buf"] = n(env.critic_obs_buf if hasattr(env, "critic_obs_buf") else env.privileged_obs_buf)
        for k, v in o.items():
            put(t, "out/" + k, v)
    for k, v in per.items():
        rec[k] = np.stack(v)
    name = args.name or f"{args.task}_n{N}"
    if spec.heightfield:
        # the terrain itself is a fixture of its own (hcr_genesis_lr_cl_b200/assets/go2_rough_terrain.npz); keep a checksum here
        rec["terrain_crc"] = np.int64(int(np.asarray(sim._height_samples.numpy(), np.int64).sum()))
    rec["meta/task"] = np.array(args.task)
    rec["meta/seed"] = np.int64(spec.seed)
    rec["meta/ctrl_delay"] = np.int64(args.ctrl_delay)
    rec["meta/gait_smooth"] = np.int64(int(args.gait_smooth))
    rec["meta/reward_names"] = np.array(sorted(spec.reward_scales))
    rec["meta/reward_scales"] = np.asarray([spec.reward_scales[k] for k in sorted(spec.reward_scales)], np.float64)
    rec["meta/contact_links"] = np.array(spec.contact_state_link_names)
    rec["meta/sum_names"] = np.array(sum_names)
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, **rec)
    resets = int(rec["out/reset_buf"].sum())
    print(f"wrote {out}: {os.path.getsize(out)/1e3:.0f} kB, resets={resets}, time_outs={int(rec['out/time_out_buf'].sum())}")
    if spec.heightfield:
        tname = "tron1_rough" if spec.robot == "tron1_pf" else "go2_rough"
        tpath = os.path.join(ROOT, "hcr_genesis_lr_cl_b200", "assets", tname + "_terrain.npz")
        if not os.path.exists(tpath):
            np.savez_compressed(tpath, height_samples=sim._height_samples.numpy().astype(np.int16),
                                terrain_origins=sim._terrain_origins.numpy().astype(np.float32))
            print(f"wrote {tpath}: {os.path.getsize(tpath)/1e3:.0f} kB")


if __name__ == "__main__":
    main()
