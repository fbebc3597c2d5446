"""Algorithmic HBM bytes per env per policy step, derived from the task descriptor.

Accounting rules: every distinct per-env tensor a kernel must read counts once for the read and
every tensor it must produce counts once for the write, per policy step; terrain gathers count the int16 bytes
touched; random numbers, constants and the task/model descriptors count zero.  `tools/algorithmic_bytes.py` prints
the table; bench.py computes `roofline.achieved` from `env_path_bytes`.  The frame stacks are double-written rings
handed out as strided views (include/b200_step.h), so a step writes each new frame twice and moves nothing else: the
"ring" variant SURVEY 8d asks to be named (its shift read of 1 563 floats and the 1 785-float stack re-write are gone;
10.4 KB per env and step in SURVEY's hand count with materialised outputs, 5.5 KB here without).
"""
from __future__ import annotations

from . import task_spec as T
from .robot_model import RobotModel


def env_kernel_items(spec: T.TaskSpec, model: RobotModel):
    A, L = spec.num_actions, model.nlinks
    feet, pen, term, cs = spec.link_groups(model)
    F, P = len(feet), spec.num_height_points if (spec.measure_heights and spec.heightfield) else 0
    w = spec.obs_widths(model)
    nsum = len(spec.episode_sum_names())
    f = 4
    reads = {
        "base pos/quat/lin/ang (world)": (3 + 4 + 3 + 3) * f, "env_origins": 3 * f, "commands": 4 * f,
        "episode_length, fail_buf": 2 * 4, "dof_pos, dof_vel": 2 * A * f, "actions, last, llast": 3 * A * f,
        "torques, last_dof_vel": 2 * A * f, "feet_pos, feet_vel": 6 * F * f, "link_contact_forces": 3 * L * f,
        "episode_sums": nsum * f, "feet_air_time, last_contacts": F * f + F, "rand_push_vels": 2 * f,
    }
    writes = {
        "base_lin_vel, base_ang_vel, projected_gravity": 9 * f, "base_quat, base_euler": 7 * f,
        "commands": 4 * f, "episode_length, fail_buf": 2 * 4, "episode_sums": nsum * f,
        "feet_air_time, last_contacts": F * f + F, "rew, reset, time_out": f + 2, "obs_buf": w["obs"] * f,
    }
    if spec.obtain_link_contact_states:
        writes["link_contact_states"] = len(cs) * f
    if P:
        reads["height scan gathers (int16 x3 per point)"] = 3 * 2 * P
        writes["measured_heights"] = P * f
        if spec.obtain_terrain_info_around_feet:
            reads["feet terrain gathers (int16 x9 per foot)"] = 9 * 2 * F
            writes["height_around_feet, normals"] = (9 + 3) * F * f
    if spec.terrain_curriculum:
        reads["terrain_levels, terrain_types (int64)"] = 16
    if w["hist"]:                                   # tasks with frame stacks: the env kernel appends one frame to each
        reads["DR params (friction, mass, com, kp, kd scales)"] = (5 + 2 * A) * f
        if w["priv"]:
            writes["privileged_obs_buf"] = w["priv"] * f
        writes["obs_history (new frame, both ring slots)"] = 2 * w["obs"] * f
        writes["critic_obs_buf (new frame, both ring slots)"] = 2 * w["single_critic"] * f
    if spec.gait_enabled:
        reads["gait_state"] = 20 * f
        writes["gait_state"] = 20 * f
    if spec.obs_kind == "go2_dreamwaq":
        writes["next_state_buf"] = w["obs"] * f
    return reads, writes


def env_kernel_bytes(spec: T.TaskSpec, model: RobotModel) -> int:
    r, w = env_kernel_items(spec, model)
    return int(sum(r.values()) + sum(w.values()))


def env_path_bytes(spec: T.TaskSpec, model: RobotModel) -> int:
    """The reference's post_physics_step = env_post_step_kernel."""
    return env_kernel_bytes(spec, model)


def shifted_stack_bytes(spec: T.TaskSpec, model: RobotModel) -> int:
    """What moving the K - 1 kept frames of both stacks every step would cost (read + write): the traffic the ring layout
    removed (round 1's history_shift_kernel), kept for the comparison in DESIGN.md."""
    w = spec.obs_widths(model)
    return 2 * 4 * (max(w["hist"] - w["obs"], 0) + max(w["critic"] - w["single_critic"], 0))


def dynamics_kernel_bytes(spec: T.TaskSpec, model: RobotModel) -> int:
    A, L = spec.num_actions, model.nlinks
    F = len(spec.link_groups(model)[0])
    f = 4
    reads = (13 + 2 * A) * f + A * f + (5 + 2 * A + 3) * f + 2 * A * f + (3 * F + 6 + A) * f   # state, actions, DR, action history, last_* sources
    writes = (13 + 2 * A) * f + A * f + 3 * L * f + 6 * F * f + 3 * A * f + (A + 3 * F + 6) * f  # state, torques, forces, feet, action history, last_*
    return int(reads + writes)


def step_bytes(spec: T.TaskSpec, model: RobotModel) -> int:
    return env_path_bytes(spec, model) + dynamics_kernel_bytes(spec, model)
