"""ctypes binding of the C ABI declared in include/b200_step.h.

The header is the single source of truth: the descriptor enums (TF_*, TI_*, RW_*, SITE_*, PHASE_*), the
B200_* limits and the field order of ``B200Buffers`` are parsed from it at import time, so the Python host
layer cannot drift from the library.  ``load_library()`` fails loudly when libb200step.so is missing -- there
is no CPU fallback (north_star).
"""
from __future__ import annotations

import ctypes
import os
import re
from collections import OrderedDict
from typing import Callable, Dict, Tuple

import numpy as np

from . import task_spec as T
from .robot_model import RobotModel

_PKG = os.path.dirname(os.path.abspath(__file__))
_HEADER_CANDIDATES = [os.path.join(os.path.dirname(_PKG), "include", "b200_step.h"), os.path.join(_PKG, "include", "b200_step.h")]
LIB_NAME = "libb200step.so"


def _parse_header():
    path = next(p for p in _HEADER_CANDIDATES if os.path.exists(p))
    src = open(path).read()
    src_nc = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    ns: Dict[str, int] = {}
    for m in re.finditer(r"#define\s+(B200_\w+)\s+(\d+)", src_nc):
        ns[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"enum\s+\w+\s*\{(.*?)\};", src_nc, flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, expr = (x.strip() for x in item.split("=", 1))
                nxt = int(eval(expr, {}, ns))  # noqa: S307 - arithmetic over names defined in the header itself
            else:
                name = item
            ns[name] = nxt
            nxt += 1
    body = re.search(r"typedef struct B200Buffers \{(.*?)\} B200Buffers;", src_nc, flags=re.S).group(1)
    fields = []
    for m in re.finditer(r"(float|int32_t|uint8_t|int64_t)\s*\*\s*(\w+)(\[2\])?\s*;", body):
        ctype, name, arr = m.groups()
        if arr:
            fields += [(name + "0", ctype), (name + "1", ctype)]
        else:
            fields.append((name, ctype))
    return ns, fields


H, BUFFER_FIELDS = _parse_header()
STATS_RING = 32      # ENV_STATS_RING in csrc/env_kernel.cuh; load_library() checks it against b200_stats_ring()


def stats_base(nsum: int) -> int:
    """Offset of the statistics ring inside `stats` (work area: sums, reset count, level sum, cstr_prob sum + padding)."""
    return 2 * max(nsum, 1) + 4
_NP = {"float": np.float32, "int32_t": np.int32, "uint8_t": np.uint8, "int64_t": np.int64}


class B200Buffers(ctypes.Structure):
    _fields_ = [(name, ctypes.c_void_p) for name, _ in BUFFER_FIELDS]


class B200RolloutTargets(ctypes.Structure):
    """include/b200_step.h: extra destinations of the next fused post step (rollout-side fusion)."""
    _fields_ = [("obs", ctypes.c_void_p), ("privileged_obs", ctypes.c_void_p), ("next_state", ctypes.c_void_p), ("rewards", ctypes.c_void_p), ("dones", ctypes.c_void_p),
                ("values", ctypes.c_void_p), ("gamma", ctypes.c_float), ("ep_return", ctypes.c_void_p), ("ep_length", ctypes.c_void_p),
                ("ep_stats", ctypes.c_void_p)]


# ------------------------------------------------------------------ descriptor packing
def pack_task(spec: T.TaskSpec, model: RobotModel, num_envs: int, hf_shape=(0, 0), env_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """TaskSpec -> (float[TF_COUNT], int[TI_COUNT])."""
    f = np.zeros(H["TF_COUNT"], np.float32)
    i = np.zeros(H["TI_COUNT"], np.int32)
    feet, pen, term, cs = spec.link_groups(model)
    w = spec.obs_widths(model)
    A = spec.num_actions
    if model.chain_len != 3:
        raise ValueError("only chains of 3 revolute joints are supported")
    if len(spec.episode_sum_names()) > 32:
        raise ValueError("more than 32 episode-sum columns")
    if len(feet) > H["B200_MAX_FEET"] or model.nlinks > H["B200_MAX_LINKS"] or w["obs"] > H["B200_MAX_OBS"]:
        raise ValueError("robot exceeds the compiled limits in b200_step.h")

    def sf(name, v, k=0):
        f[H[name] + k] = v

    def span(lo_name, span_name, rng):
        f[H[lo_name]] = rng[0]
        f[H[span_name]] = np.float32(float(rng[1]) - float(rng[0]))

    sf("TF_SIM_DT", spec.sim_dt); sf("TF_POLICY_DT", spec.dt); sf("TF_ACTION_SCALE", spec.action_scale)
    sf("TF_KP", spec.kp); sf("TF_KD", spec.kd); sf("TF_CLIP_ACTIONS", spec.clip_actions); sf("TF_CLIP_OBS", spec.clip_observations)
    for k in range(3):
        sf("TF_INIT_POS", spec.init_pos[k], k)
    for k in range(4):
        sf("TF_INIT_QUAT", spec.init_quat_xyzw[k], k)
    sf("TF_RESET_ROOT_XY", spec.reset_root_xy); sf("TF_RESET_ROOT_VEL", spec.reset_root_vel)
    sf("TF_FAIL_LIMIT", spec.fail_limit); sf("TF_MAX_PROJ_GRAV", spec.max_projected_gravity)
    sf("TF_HSCALE", spec.horizontal_scale); sf("TF_VSCALE", spec.vertical_scale)
    sf("TF_BORDER", spec.border_size if spec.heightfield else 0.0)
    if spec.heightfield:        # genesis_simulator.py:278-294
        sf("TF_X_LO", -spec.border_size + 1.0); sf("TF_X_HI", spec.border_size + spec.num_rows * spec.terrain_length - 1.0)
        sf("TF_Y_LO", -spec.border_size + 1.0); sf("TF_Y_HI", spec.border_size + spec.num_cols * spec.terrain_width - 1.0)
    else:
        sf("TF_X_LO", -spec.plane_length / 2 + 1); sf("TF_X_HI", spec.plane_length / 2 - 1)
        sf("TF_Y_LO", -spec.plane_length / 2 + 1); sf("TF_Y_HI", spec.plane_length / 2 - 1)
    sf("TF_TERRAIN_HALF_LENGTH", spec.terrain_length / 2); sf("TF_EPISODE_LENGTH_S", spec.episode_length_s)
    span("TF_CMD_VY_LO", "TF_CMD_VY_SPAN", spec.cmd_lin_vel_y)
    span("TF_CMD_YAW_LO", "TF_CMD_YAW_SPAN", spec.cmd_ang_vel_yaw); sf("TF_CMD_YAW_HI", spec.cmd_ang_vel_yaw[1])
    span("TF_CMD_HEADING_LO", "TF_CMD_HEADING_SPAN", spec.cmd_heading)
    span("TF_FRICTION_LO", "TF_FRICTION_SPAN", spec.friction_range); span("TF_MASS_LO", "TF_MASS_SPAN", spec.added_mass_range)
    span("TF_COMX_LO", "TF_COMX_SPAN", spec.com_pos_x_range); span("TF_COMY_LO", "TF_COMY_SPAN", spec.com_pos_y_range)
    span("TF_COMZ_LO", "TF_COMZ_SPAN", spec.com_pos_z_range)
    span("TF_KPS_LO", "TF_KPS_SPAN", spec.kp_range); span("TF_KDS_LO", "TF_KDS_SPAN", spec.kd_range)
    span("TF_ARM_LO", "TF_ARM_SPAN", spec.joint_armature_range); span("TF_JFR_LO", "TF_JFR_SPAN", spec.joint_friction_range)
    span("TF_JDA_LO", "TF_JDA_SPAN", spec.joint_damping_range)
    sf("TF_MAX_PUSH", spec.max_push_vel_xy)
    sf("TF_FRICTION_OFFSET", (spec.friction_range[0] + spec.friction_range[1]) / 2)   # legged_robot.py:448-453
    sf("TF_KPS_OFFSET", (spec.kp_range[0] + spec.kp_range[1]) / 2); sf("TF_KDS_OFFSET", (spec.kd_range[0] + spec.kd_range[1]) / 2)
    sf("TF_OS_LIN_VEL", spec.obs_scale_lin_vel); sf("TF_OS_ANG_VEL", spec.obs_scale_ang_vel)
    sf("TF_OS_DOF_POS", spec.obs_scale_dof_pos); sf("TF_OS_DOF_VEL", spec.obs_scale_dof_vel)
    sf("TF_OS_HEIGHT", spec.obs_scale_height); sf("TF_HEIGHT_OBS_OFFSET", spec.height_obs_offset)
    sf("TF_TRACKING_SIGMA", spec.tracking_sigma); sf("TF_BASE_HEIGHT_TARGET", spec.base_height_target)
    sf("TF_FOOT_CLEARANCE_TARGET", spec.foot_clearance_target); sf("TF_FOOT_HEIGHT_OFFSET", spec.foot_height_offset)
    sf("TF_FOOT_CLEARANCE_SIGMA", spec.foot_clearance_tracking_sigma); sf("TF_ABOUT_LANDING", spec.about_landing_threshold)
    sf("TF_AIR_TIME_THRESHOLD", spec.feet_air_time_threshold); sf("TF_FOOT_DISTANCE_THRESHOLD", spec.foot_distance_threshold)
    # physics formulation constants: MuJoCo-style soft constraints
    sf("TF_GRAV", 9.81); sf("TF_TC", 2 * spec.sim_dt); sf("TF_DAMPRATIO", 1.0)
    sf("TF_D0", 0.9); sf("TF_DMAX", 0.95); sf("TF_WIDTH", 0.001); sf("TF_MID", 0.5); sf("TF_POWER", 2.0)
    sf("TF_TERRAIN_MU", spec.static_friction); sf("TF_GEOM_MU", 1.0); sf("TF_PGS_TOL", spec.pgs_tolerance)
    sf("TF_GAIT_PERIOD", spec.gait_period); sf("TF_GAIT_THETA_LEFT", spec.gait_theta_left)
    sf("TF_GAIT_THETA_RIGHT", np.float32(spec.gait_theta_right - spec.gait_theta_left))      # stored as the offset right - left
    sf("TF_GAIT_B_SWING", np.float32(spec.gait_b_swing * 2) * np.float32(np.pi))            # b_swing * 2 * torch.pi in fp32
    sf("TF_BASE_HEIGHT_SIGMA", spec.base_height_tracking_sigma); sf("TF_SIT_PERCENT", spec.sit_init_percent)
    for k in range(3):
        sf("TF_SIT_POS", spec.sit_pos[k], k)
    half = np.float32(spec.sit_pitch_angle) * np.float32(0.5)
    for k, v in enumerate((0.0, np.sin(half), 0.0, np.cos(half))):                          # quat_from_euler_xyz(0, pitch, 0), xyzw
        sf("TF_SIT_QUAT", v, k)
    for j, v in enumerate(spec.sit_joint_angles):
        sf("TF_SIT_DOF_POS", v, j)
    sf("TF_EULER_SIGMA", spec.euler_tracking_sigma)
    for gi, th in enumerate(spec.gait_theta_lists[:H["B200_MAX_GAITS"]]):
        for fi, v in enumerate(th):
            sf("TF_GAIT_THETA", v, gi * H["B200_MAX_FEET"] + fi)
    sf("TF_CAT_SOFT_P", spec.cat_soft_p); sf("TF_CAT_ACTION_RATE", spec.cat_action_rate)
    sf("TF_CAT_MIN_BASE_HEIGHT", spec.cat_min_base_height); sf("TF_CAT_MAX_PROJ_GRAV", spec.cat_max_projected_gravity)
    lim = soft_dof_limits(spec, model)
    for j in range(A):
        sf("TF_DEFAULT_DOF_POS", spec.default_dof_pos[j], j); sf("TF_RESET_DOF_NOISE", spec.reset_dof_noise[j], j)
        sf("TF_DOF_LIM_LO", lim[j, 0], j); sf("TF_DOF_LIM_HI", lim[j, 1], j)
        sf("TF_TORQUE_LIMIT", model.effort[j], j)
        sf("TF_DOF_VEL_LIMIT", spec.dof_vel_limits[j] if spec.dof_vel_limits else model.velocity[j], j)
    for name in T.REWARD_TERMS:
        sf("TF_REWARD_SCALE", spec.scaled_reward(name), T.REWARD_ID[name])
    nv = spec.noise_scale_vec()
    f[H["TF_NOISE_VEC"]:H["TF_NOISE_VEC"] + len(nv)] = nv
    px, py = spec.measured_points_x, spec.measured_points_y
    if len(px) > H["B200_MAX_PTS_AXIS"] or len(py) > H["B200_MAX_PTS_AXIS"]:
        raise ValueError("too many height-scan points per axis")
    f[H["TF_POINTS_X"]:H["TF_POINTS_X"] + len(px)] = px
    f[H["TF_POINTS_Y"]:H["TF_POINTS_Y"] + len(py)] = py

    def si(name, v, k=0):
        i[H[name] + k] = int(v)

    si("TI_NUM_ENVS", num_envs); si("TI_A", A); si("TI_C", model.num_chains); si("TI_D", model.chain_len)
    si("TI_L", model.nlinks); si("TI_F", len(feet)); si("TI_PX", len(px)); si("TI_PY", len(py)); si("TI_NSPHERES", model.nspheres)
    si("TI_DECIMATION", spec.decimation); si("TI_PGS_ITERS", spec.pgs_iterations)
    si("TI_OBS_KIND", T.OBS_KINDS[spec.obs_kind]); si("TI_NUM_OBS", w["obs"]); si("TI_NUM_PRIV", w["priv"])
    si("TI_SINGLE_CRITIC", w["single_critic"]); si("TI_FRAME_STACK", spec.frame_stack); si("TI_C_FRAME_STACK", spec.c_frame_stack)
    si("TI_MAX_EPISODE_LENGTH", spec.max_episode_length); si("TI_RESAMPLE_INTERVAL", spec.resample_interval)
    si("TI_PUSH_INTERVAL", spec.push_interval)
    si("TI_TRIMESH", spec.mesh_type == "trimesh")
    si("TI_HEIGHTFIELD", spec.heightfield); si("TI_MEASURE_HEIGHTS", spec.measure_heights and spec.heightfield)
    si("TI_FEET_INFO", spec.obtain_terrain_info_around_feet and spec.measure_heights and spec.heightfield)
    si("TI_TERRAIN_CURRICULUM", spec.terrain_curriculum and spec.heightfield); si("TI_HEADING_COMMAND", spec.heading_command)
    si("TI_PUSH_ROBOTS", spec.push_robots); si("TI_ADD_NOISE", spec.add_noise); si("TI_ONLY_POSITIVE", spec.only_positive_rewards)
    si("TI_RAND_FRICTION", spec.randomize_friction); si("TI_RAND_MASS", spec.randomize_base_mass)
    si("TI_RAND_COM", spec.randomize_com_displacement); si("TI_RAND_PD", spec.randomize_pd_gain)
    si("TI_RAND_ARMATURE", spec.randomize_joint_armature); si("TI_RAND_JFRICTION", spec.randomize_joint_friction)
    si("TI_RAND_JDAMPING", spec.randomize_joint_damping); si("TI_CONTACT_STATES", spec.obtain_link_contact_states)
    si("TI_CLEARANCE_USES_TERRAIN", spec.foot_clearance_uses_terrain)
    si("TI_NUM_LEVELS", spec.num_rows); si("TI_NUM_TYPES", spec.num_cols); si("TI_HF_ROWS", hf_shape[0]); si("TI_HF_COLS", hf_shape[1])
    si("TI_SEED_LO", spec.seed & 0x7FFFFFFF); si("TI_SEED_HI", (spec.seed >> 32) & 0x7FFFFFFF)
    active = spec.active_rewards()
    si("TI_N_REWARDS", len(active))
    for k, name in enumerate(active):
        si("TI_REWARD_IDS", T.REWARD_ID[name], k)
    sums = spec.episode_sum_names()
    si("TI_TERMINATION_COL", sums.index("termination") if "termination" in sums else -1)
    si("TI_N_PEN", len(pen)); si("TI_N_TERM", len(term)); si("TI_N_CS", len(cs)); si("TI_ENV_OFFSET", env_offset)
    si("TI_CAT", spec.cat_enabled); si("TI_CAT_GLOBAL_STANDSTILL", spec.cat_stand_still_global)
    si("TI_DOUBLE_SHIFT", spec.double_shift_actions); si("TI_N_SUMS", len(spec.episode_sum_names()))
    si("TI_GAIT", spec.gait_enabled); si("TI_CLEARANCE_MODE", spec.foot_clearance_mode)
    si("TI_R18", spec.gait_enabled and spec.reproduce_r18)
    if spec.gait_enabled and spec.gait_smooth:
        p_terms, tab = spec.von_mises_series()
        if p_terms > H["B200_VM_MAX"]:
            raise ValueError("von Mises series longer than B200_VM_MAX")
        si("TI_VM_TERMS", p_terms)
        f[H["TF_VM_R"]:H["TF_VM_R"] + len(tab)] = tab
    si("TI_NUM_TEACHER", spec.num_teacher)
    if spec.randomize_ctrl_delay:
        lo, hi = (int(v) for v in spec.ctrl_delay_step_range)
        if not 0 <= lo <= hi <= H["B200_MAX_CTRL_DELAY"]:
            raise ValueError("ctrl_delay_step_range outside [0, B200_MAX_CTRL_DELAY]")
        si("TI_CTRL_DELAY", 1); si("TI_CTRL_DELAY_LO", lo); si("TI_CTRL_DELAY_HI", hi)
    si("TI_BEHAVIOR", spec.behavior_enabled); si("TI_BEHAVIOR_INTERVAL", max(int(spec.behavior_resampling_time / spec.dt), 1))
    for k, v in enumerate(feet):
        si("TI_FEET_LINKS", v, k)
    for k, v in enumerate(pen):
        si("TI_PEN_LINKS", v, k)
    for k, v in enumerate(term):
        si("TI_TERM_LINKS", v, k)
    for k, v in enumerate(cs):
        si("TI_CS_LINKS", v, k)
    return f, i


def soft_dof_limits(spec: T.TaskSpec, model: RobotModel) -> np.ndarray:
    """genesis_simulator.py:366-382 in fp32."""
    f32 = np.float32
    lim = model.dof_limits.astype(np.float32).copy()
    for j in range(lim.shape[0]):
        m = (lim[j, 0] + lim[j, 1]) / f32(2)
        r = lim[j, 1] - lim[j, 0]
        lim[j, 0] = m - f32(0.5) * r * f32(spec.soft_dof_pos_limit)
        lim[j, 1] = m + f32(0.5) * r * f32(spec.soft_dof_pos_limit)
    return lim


def buffer_shapes(spec: T.TaskSpec, model: RobotModel, N: int) -> "OrderedDict[str, Tuple[tuple, type]]":
    """Shape and dtype of every B200Buffers field, in ABI order."""
    A, L = spec.num_actions, model.nlinks
    feet, _, _, cs = spec.link_groups(model)
    F, P = len(feet), max(spec.num_height_points, 1)
    w = spec.obs_widths(model)
    nsum = len(spec.episode_sum_names())
    shp = dict(
        base_pos=(N, 3), base_quat_wxyz=(N, 4), base_lin_w=(N, 3), base_ang_w=(N, 3), dof_pos=(N, A), dof_vel=(N, A),
        friction=(N, 1), added_mass=(N, 1), com_bias=(N, 3), kp_scale=(N, A), kd_scale=(N, A), joint_armature=(N, 1),
        joint_friction=(N, 1), joint_damping=(N, 1), rand_push_vels=(N, 3),
        actions=(N, A), last_actions=(N, A), llast_actions=(N, A), commands=(N, 4), episode_length=(N,), fail_buf=(N,),
        feet_air_time=(N, F), last_contacts=(N, F), episode_sums=(N, max(nsum, 1)), terrain_levels=(N,), terrain_types=(N,),
        env_origins=(N, 3),
        base_quat=(N, 4), base_euler=(N, 3), base_lin_vel=(N, 3), base_ang_vel=(N, 3), projected_gravity=(N, 3),
        torques=(N, A), link_contact_forces=(N, L, 3), feet_pos=(N, F, 3), feet_vel=(N, F, 3),
        link_contact_states=(N, max(len(cs), 1)), measured_heights=(N, P), height_around_feet=(N, F, 9),
        normal_vector_around_feet=(N, 3 * F), last_dof_vel=(N, A), last_feet_vel=(N, F, 3), last_base_lin_vel=(N, 3),
        last_base_ang_vel=(N, 3),
        obs_buf=(N, w["obs"]), privileged_obs_buf=(N, max(w["priv"], 1)),
        # double-written rings of period K + 1 (b200_step.h)
        obs_history=(N, 2 * (spec.frame_stack + 1) * w["obs"] if w["hist"] else 2), critic_obs=(N, 2 * (spec.c_frame_stack + 1) * w["single_critic"] if w["critic"] else 2),
        rew_buf=(N,), reset_buf=(N,), time_out_buf=(N,), contact_warm=(N, 48), gait_state=(N, H["B200_GAIT_STATE"]), height_cells=(N, P, 2), stats=(stats_base(nsum) + STATS_RING * (max(nsum, 1) + H["B200_STATS_EXTRA"]),), cstr_prob=(N,), global_flags=(4,),
        next_state_buf=(N, w["obs"] if spec.obs_kind == "go2_dreamwaq" else 1), dyn_cost=(2, N), dyn_order=(2, N),
        action_queue=(N, (int(spec.ctrl_delay_step_range[1]) + 1) * A if spec.randomize_ctrl_delay else 1), action_delay=(N,),
        nonfinite=(N,),
    )
    return OrderedDict((name, (shp[name], _NP[ct])) for name, ct in BUFFER_FIELDS)


def fill_buffers(ptr_of: Callable[[str], int]) -> B200Buffers:
    b = B200Buffers()
    for name, _ in BUFFER_FIELDS:
        setattr(b, name, ptr_of(name))
    return b


# ------------------------------------------------------------------ the library
_lib = None


def lib_path() -> str:
    return os.path.join(_PKG, LIB_NAME)


def load_library() -> ctypes.CDLL:
    """Load libb200step.so; raise if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build the CUDA extension first (python -m hcr_genesis_lr_cl_b200.build); "
                           "this backend has no CPU or PyTorch fallback")
    _lib = bind(ctypes.CDLL(path))
    return _lib


def bind(lib: ctypes.CDLL) -> ctypes.CDLL:
    """Declare the argument / result types of every entry point of include/b200_step.h on a loaded library."""
    vp, ip, fp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)
    lib.b200_create.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.POINTER(vp)]
    lib.b200_create.restype = ctypes.c_int
    lib.b200_destroy.argtypes = [vp]
    lib.b200_destroy.restype = None
    lib.b200_set_terrain.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int]
    lib.b200_set_terrain.restype = ctypes.c_int
    lib.b200_bind_buffers.argtypes = [vp, ctypes.POINTER(B200Buffers)]
    lib.b200_bind_buffers.restype = ctypes.c_int
    lib.b200_dynamics_step.argtypes = [vp, vp, vp]
    lib.b200_dynamics_step.restype = ctypes.c_int
    lib.b200_simulator_step.argtypes = [vp, vp, vp]
    lib.b200_simulator_step.restype = ctypes.c_int
    lib.b200_stats_ring.argtypes = []
    lib.b200_stats_ring.restype = ctypes.c_int
    lib.b200_device.argtypes = [vp]
    lib.b200_device.restype = ctypes.c_int
    lib.b200_env_post_step.argtypes = [vp, ctypes.c_longlong, ctypes.c_float, ctypes.c_float, ctypes.c_longlong, ctypes.c_int, vp]
    lib.b200_env_post_step.restype = ctypes.c_int
    lib.b200_env_step.argtypes = [vp, vp, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_float, ctypes.c_longlong, vp, vp, vp, vp]
    lib.b200_env_step.restype = ctypes.c_int
    lib.b200_env_step_device.argtypes = [vp, vp, vp, ctypes.c_double, vp]
    lib.b200_env_step_device.restype = ctypes.c_int
    lib.b200_set_rollout_targets.argtypes = [vp, ctypes.POINTER(B200RolloutTargets)]
    lib.b200_set_rollout_targets.restype = ctypes.c_int
    lib.b200_set_step_flags.argtypes = [vp, ctypes.c_int]
    lib.b200_set_step_flags.restype = ctypes.c_int
    lib.b200_set_side_stream.argtypes = [vp, ctypes.c_int]
    lib.b200_set_side_stream.restype = ctypes.c_int
    lib.b200_set_dynamics_order.argtypes = [vp, ctypes.c_int]
    lib.b200_set_dynamics_order.restype = ctypes.c_int
    lib.b200_set_behavior.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
    lib.b200_set_behavior.restype = ctypes.c_int
    lib.b200_reset_all.argtypes = [vp, ctypes.c_longlong, ctypes.c_float, ctypes.c_float, vp]
    lib.b200_reset_all.restype = ctypes.c_int
    lib.b200_kernel_info.argtypes = [vp, ctypes.c_char_p, ip, ip, ip, ip]
    lib.b200_kernel_info.restype = ctypes.c_int
    lib.b200_env_kernel_variant.argtypes = [vp]
    lib.b200_env_kernel_variant.restype = ctypes.c_char_p
    lib.b200_launch_count.argtypes = [vp]
    lib.b200_launch_count.restype = ctypes.c_longlong
    lib.b200_last_error.argtypes = []
    lib.b200_last_error.restype = ctypes.c_char_p
    if lib.b200_stats_ring() != STATS_RING:
        raise RuntimeError("libb200step.so and _cabi.py disagree on the extras ring length (rebuild the extension)")
    return lib


EXPORTED_SYMBOLS = ["b200_create", "b200_destroy", "b200_set_terrain", "b200_bind_buffers", "b200_dynamics_step", "b200_simulator_step",
                    "b200_stats_ring", "b200_device",
                    "b200_set_side_stream", "b200_set_dynamics_order", "b200_env_post_step", "b200_env_step", "b200_env_step_device", "b200_set_step_flags", "b200_set_rollout_targets", "b200_set_behavior", "b200_reset_all", "b200_kernel_info", "b200_env_kernel_variant", "b200_launch_count", "b200_last_error"]
