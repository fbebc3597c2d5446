"""TEST-ONLY helpers for plugin mode: the reference's own, unmodified task classes (LeggedRobot, GO2, Go2TS, GO2WTW,
Go2CaT, TRON1PF_EE, ...) stepping over ``B200Simulator`` through the ``Simulator`` plugin API, next to the fused env.

The reference tree comes from /root/reference (build container) or the staged copy baseline/_ref (GPU box,
tools/stage_reference.py); the tests skip when neither is there."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    for p in (os.environ.get("HCR_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if p and os.path.isdir(os.path.join(p, "legged_gym")):
            return p
    return None


def install_plugin(impl=None):
    """hcr_genesis_lr_cl_b200.plugin.install() on the reference tree (+ import stubs for trimesh / matplotlib / pygame /
    xlsxwriter, which legged_gym imports at module level but the path never uses)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("no reference tree")
    for p in (os.path.join(ROOT, "oracle", "stubs_thirdparty"), os.path.join(ROOT, "oracle", "stubs"), root):
        # oracle/stubs holds the stand-in `genesis` engine of the golden harness; on sys.path here only so that both
        # harnesses can live in one pytest process -- plugin mode never calls into it (see test_placeholder_genesis)
        if p not in sys.path:
            sys.path.insert(0, p)
    from hcr_genesis_lr_cl_b200 import plugin
    return plugin.install(impl=impl)


def uninstall_plugin():
    from hcr_genesis_lr_cl_b200 import plugin
    plugin.uninstall()


def make_plugin_env(task, num_envs, impl=None, cpu=True, cfg_edit=None):
    """task_registry.make_env(task) of the reference with the B200 backend behind BaseTask's dispatch."""
    backend = install_plugin(impl)
    import legged_gym.envs  # noqa: F401  (registers the tasks)
    from legged_gym.utils.task_registry import task_registry
    args = SimpleNamespace(task=task, headless=True, cpu=cpu, num_envs=num_envs, debug=False, max_iterations=None, resume=False,
                           sync_wandb=False, ckpt=-1, load_run=None, export_onnx=False, use_joystick=False, joystick_type=None,
                           follow_robot=False)
    env_cfg, train_cfg = task_registry.get_cfgs(name=task)
    if cfg_edit is not None:
        cfg_edit(env_cfg)
    env, env_cfg = task_registry.make_env(task, args=args, env_cfg=env_cfg)
    assert isinstance(env.simulator, backend), type(env.simulator)
    return env, env_cfg, train_cfg


def _n(t):
    return t.detach().cpu().numpy().copy()


def pack_gait(env):
    """theta[4], gait_time, phi, gait_period, base_height_t, foot_clearance_t, pitch_t, clock[8] (B200Buffers.gait_state)."""
    N = env.num_envs
    g = np.zeros((N, 20), np.float32)
    th, clk = _n(env.theta), _n(env.clock_input)
    g[:, 0:th.shape[1]] = th
    g[:, 4:5], g[:, 5:6], g[:, 6:7] = _n(env.gait_time), _n(env.phi), _n(env.gait_period)
    for k, name in ((7, "base_height_target"), (8, "foot_clearance_target"), (9, "pitch_target")):
        if hasattr(env, name):
            g[:, k:k + 1] = _n(getattr(env, name))
    g[:, 10:10 + clk.shape[1]] = clk
    return g


# simulator-side buffers that make up the state a fused env needs to continue a plugin-mode run
SIM_STATE = ["base_pos", "base_quat_wxyz", "base_lin_w", "base_ang_w", "dof_pos", "dof_vel", "friction", "added_mass", "com_bias",
             "kp_scale", "kd_scale", "joint_armature", "joint_friction", "joint_damping", "rand_push_vels", "terrain_levels",
             "terrain_types", "env_origins", "base_quat", "base_euler", "base_lin_vel", "base_ang_vel", "projected_gravity", "torques",
             "link_contact_forces", "feet_pos", "feet_vel", "link_contact_states", "measured_heights", "height_around_feet",
             "normal_vector_around_feet", "last_dof_vel", "last_feet_vel", "last_base_lin_vel", "last_base_ang_vel", "contact_warm"]


def transfer_state(ref_env, fused_env, sum_names):
    """Copy the whole state of a reference env running over B200Simulator (plugin mode) into a FusedLeggedEnv."""
    rs, fs = ref_env.simulator, fused_env.simulator
    fb = fs._buf
    for k in SIM_STATE:
        if k in fb and k in rs._buf and fb[k].shape == rs._buf[k].shape:
            fb[k].copy_(rs._buf[k])
    st = dict(actions=ref_env.actions, last_actions=ref_env.last_actions, llast_actions=ref_env.llast_actions, commands=ref_env.commands,
              episode_length=ref_env.episode_length_buf, fail_buf=ref_env.fail_buf, feet_air_time=ref_env.feet_air_time,
              last_contacts=ref_env.last_contacts)
    for k, v in st.items():
        fb[k].copy_(v.reshape(fb[k].shape).to(fb[k].dtype))
    fb["episode_sums"].copy_(torch.stack([ref_env.episode_sums[k] for k in sum_names], dim=1))
    if hasattr(ref_env, "gait_time"):
        fb["gait_state"].copy_(torch.from_numpy(pack_gait(ref_env)).to(fb["gait_state"].device))
    hist = crit = None
    if hasattr(ref_env, "critic_history"):                      # go2_wtw names its deques obs_history / critic_history
        hist, crit = ref_env.obs_history, ref_env.critic_history
    elif hasattr(ref_env, "obs_history_deque"):
        hist, crit = ref_env.obs_history_deque, ref_env.critic_obs_deque
    if hist is not None:
        fs.load_state(dict(obs_history=torch.cat(list(hist), dim=1).cpu().numpy(), critic_obs=torch.cat(list(crit), dim=1).cpu().numpy()))
    if getattr(ref_env.cfg.domain_rand, "randomize_ctrl_delay", False):
        fb["action_queue"].copy_(ref_env.action_queue.reshape(fb["action_queue"].shape))
        fb["action_delay"].copy_(ref_env.action_delay.to(torch.int32))
    fused_env.common_step_counter = int(ref_env.common_step_counter)
    fused_env.command_ranges["lin_vel_x"] = [float(x) for x in ref_env.command_ranges["lin_vel_x"]]
    if hasattr(ref_env, "gait_period_range"):                   # go2_wtw host-side behaviour curriculum state
        fused_env.gait_period_range, fused_env.base_height_target_range = list(ref_env.gait_period_range), list(ref_env.base_height_target_range)
        fused_env.foot_clearance_target_range, fused_env.pitch_target_range = list(ref_env.foot_clearance_target_range), list(ref_env.pitch_target_range)
        fused_env.num_gaits = int(ref_env.num_gaits)
