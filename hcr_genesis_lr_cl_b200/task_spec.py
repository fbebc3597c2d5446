"""Task descriptor for the fused environment step.

A ``TaskSpec`` is the flat, numbers-only description of one reference task that the CUDA
kernels need: robot, control gains, terrain, command ranges, domain randomisation ranges, reward
scales and the observation layout.  It mirrors (values only) what the reference spreads over its
nested config classes:

* ``LeggedRobotCfg``             legged_gym/envs/base/legged_robot_config.py:3-272
* ``Go2FlatCommonCfg`` / ``Go2RoughCommonCfg``  legged_gym/envs/base/common_cfgs.py:9-129
* ``GO2Cfg``                     legged_gym/envs/go2/go2_config.py:5-73
* ``Go2TSCfg``                   legged_gym/envs/go2/go2_ts/go2_ts_config.py:5-67

``TaskSpec.from_reference_cfg(cfg, task)`` builds the same object from a live reference config
(what the ``B200Simulator`` drop-in does inside LeggedGym-Ex); ``go2_spec()`` / ``go2_ts_spec()``
are the built-in presets used when the reference tree is absent (GPU box).  tests/test_task_spec.py
checks the presets against the reference configs whenever /root/reference is present.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from .robot_model import GO2_DOF_NAMES, TRON1_PF_DOF_NAMES, RobotModel, load_robot_model

# Reward terms, alphabetical: the reference evaluates `_reward_<name>` in the order of
# class_to_dict(cfg.rewards.scales) == dir() order (legged_gym/utils/helpers.py:10-25,
# legged_robot.py:411-434); "termination" is always added last (legged_robot.py:163-168).
REWARD_TERMS: List[str] = [
    "action_rate", "action_smoothness", "ang_vel_xy", "base_height", "biped_periodic_gait", "collision", "dof_acc",
    "dof_close_to_default", "dof_pos_limits", "dof_pos_stand_still", "dof_power", "dof_vel",
    "dof_vel_stand_still", "feet_air_time", "feet_contact_stand_still", "feet_distance", "foot_acc", "foot_clearance",
    "foot_landing_vel", "hip_pos", "keep_balance", "lin_vel_z", "no_fly", "orientation", "quad_periodic_gait", "thigh_pos", "torques",
    "tracking_ang_vel", "tracking_base_height", "tracking_foot_clearance", "tracking_lin_vel", "tracking_orientation", "termination",
]
REWARD_ID: Dict[str, int] = {n: i for i, n in enumerate(REWARD_TERMS)}
NUM_REWARD_TERMS = len(REWARD_TERMS)

#: observation layouts the fused kernel knows (per-task ``compute_observations``)
OBS_KINDS = {"go2": 0, "go2_ts": 1, "go2_cat": 2, "tron1_pf": 3, "tron1_pf_ee": 4, "go2_wtw": 5, "go2_cts": 6, "go2_ee": 7,
             "go2_dreamwaq": 8}

#: `task` of the descriptor B200Simulator derives in plugin mode (no task name is handed to a Simulator backend,
#: base_task.py:41-48): physics, terrain, DR and link groups only -- rewards / observations stay the task class's own code
PLUGIN_TASK = "plugin"

#: reference config class -> registered task name (legged_gym/envs/__init__.py:80-93); lets the fused registration of
#: INTEGRATION.md derive its descriptor from the cfg object alone
CFG_CLASS_TO_TASK = {"GO2Cfg": "go2", "GO2WTWCfg": "go2_wtw", "Go2TSCfg": "go2_ts", "Go2EECfg": "go2_ee", "Go2CTSCfg": "go2_cts",
                     "Go2DreamwaqCfg": "go2_dreamwaq", "Go2CaTCfg": "go2_cat", "TRON1PFCfg": "tron1_pf", "TRON1PF_EECfg": "tron1_pf_ee"}


def task_of_cfg(cfg) -> Optional[str]:
    """Registered task name of a reference config object (most derived known class wins), or None."""
    for klass in type(cfg).__mro__:
        if klass.__name__ in CFG_CLASS_TO_TASK:
            return CFG_CLASS_TO_TASK[klass.__name__]
    return None


CAT_CONSTRAINTS = ["torque", "dof_vel", "action_rate", "base_height", "collision", "feet_stumble", "dof_pos",
                   "base_orientation", "stand_still"]      # order of ConstraintManager.add calls, go2_cat.py:197-208

# Philox draw sites (see oracle/philox.py and csrc/philox.cuh)
SITE_CMD_RESAMPLE, SITE_PUSH, SITE_LEVEL, SITE_CMD_RESET, SITE_DOF, SITE_ROOT, SITE_FRICTION, SITE_MASS, SITE_COM, \
    SITE_KP, SITE_KD, SITE_ARMATURE, SITE_JFRICTION, SITE_JDAMPING, SITE_OBS_NOISE, SITE_GAIT, SITE_HOST, SITE_BEHAVIOR, SITE_BEHAVIOR_RESET, \
    SITE_CTRL_DELAY = range(20)


@dataclass
class TaskSpec:
    task: str
    obs_kind: str
    robot: str
    dof_names: List[str]
    # sizes
    num_obs: int
    num_privileged_obs: Optional[int]
    frame_stack: int = 1
    c_frame_stack: int = 1
    # control (common_cfgs.py:33-41)
    sim_dt: float = 0.005
    decimation: int = 4
    action_scale: float = 0.25
    kp: float = 20.0
    kd: float = 0.5
    default_dof_pos: List[float] = field(default_factory=list)
    clip_actions: float = 100.0
    clip_observations: float = 100.0
    # init state
    init_pos: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.42])
    init_quat_xyzw: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0, 1.0])
    reset_dof_noise: List[float] = field(default_factory=list)   # per-dof half range of U(-r, r) added to q0
    reset_root_xy: float = 0.5        # U(+-) on xy when custom origins (legged_robot.py:288)
    reset_root_vel: float = 0.5       # U(+-) on lin/ang vel (legged_robot.py:295-297; go2.py:131-133 uses 0)
    # episode
    episode_length_s: float = 20.0
    fail_to_terminal_time_s: float = 0.1
    max_projected_gravity: float = -0.1
    send_timeouts: bool = True
    # terrain (legged_robot_config.py:19-51, common_cfgs.py:75-93)
    mesh_type: str = "plane"
    horizontal_scale: float = 0.1
    vertical_scale: float = 0.005
    border_size: float = 5.0
    plane_length: float = 200.0
    env_spacing: float = 1.0
    terrain_length: float = 6.0
    terrain_width: float = 6.0
    num_rows: int = 4
    num_cols: int = 4
    terrain_curriculum: bool = False
    max_init_terrain_level: int = 1
    static_friction: float = 1.0
    measure_heights: bool = False
    measured_points_x: List[float] = field(default_factory=list)
    measured_points_y: List[float] = field(default_factory=list)
    obtain_terrain_info_around_feet: bool = False
    # commands (legged_robot_config.py:131-142)
    num_commands: int = 4
    resampling_time: float = 10.0
    heading_command: bool = True
    cmd_curriculum: bool = False
    max_curriculum: float = 1.0
    curriculum_threshold: float = 0.8
    cmd_lin_vel_x: List[float] = field(default_factory=lambda: [-1.0, 1.0])
    cmd_lin_vel_y: List[float] = field(default_factory=lambda: [-1.0, 1.0])
    cmd_ang_vel_yaw: List[float] = field(default_factory=lambda: [-1.0, 1.0])
    cmd_heading: List[float] = field(default_factory=lambda: [-3.14, 3.14])
    # domain randomisation (legged_robot_config.py:144-177)
    randomize_friction: bool = True
    friction_range: List[float] = field(default_factory=lambda: [0.5, 1.25])
    randomize_base_mass: bool = True
    added_mass_range: List[float] = field(default_factory=lambda: [-1.0, 1.0])
    push_robots: bool = True
    push_interval_s: float = 15.0
    max_push_vel_xy: float = 1.0
    randomize_com_displacement: bool = True
    com_pos_x_range: List[float] = field(default_factory=lambda: [-0.01, 0.01])
    com_pos_y_range: List[float] = field(default_factory=lambda: [-0.01, 0.01])
    com_pos_z_range: List[float] = field(default_factory=lambda: [-0.01, 0.01])
    randomize_pd_gain: bool = False
    kp_range: List[float] = field(default_factory=lambda: [0.8, 1.2])
    kd_range: List[float] = field(default_factory=lambda: [0.8, 1.2])
    randomize_joint_armature: bool = False
    joint_armature_range: List[float] = field(default_factory=lambda: [0.0, 0.05])
    randomize_joint_friction: bool = False
    joint_friction_range: List[float] = field(default_factory=lambda: [0.0, 0.1])
    randomize_joint_damping: bool = False
    joint_damping_range: List[float] = field(default_factory=lambda: [0.0, 1.0])
    # normalisation / noise (legged_robot_config.py:179-201)
    obs_scale_lin_vel: float = 1.0
    obs_scale_ang_vel: float = 0.25
    obs_scale_dof_pos: float = 1.0
    obs_scale_dof_vel: float = 0.05
    obs_scale_height: float = 5.0
    add_noise: bool = True
    noise_level: float = 1.0
    noise_dof_pos: float = 0.01
    noise_dof_vel: float = 0.5
    noise_lin_vel: float = 0.1
    noise_ang_vel: float = 0.2
    noise_gravity: float = 0.05
    noise_height: float = 0.1
    height_obs_offset: float = 0.5     # clip(base_z - 0.5 - h, -1, 1) (go2_ts.py:45-46)
    # rewards (legged_robot_config.py:105-129 + task overrides)
    reward_scales: Dict[str, float] = field(default_factory=dict)  # as in cfg (NOT yet * dt)
    only_positive_rewards: bool = True
    tracking_sigma: float = 0.25
    soft_dof_pos_limit: float = 1.0
    base_height_target: float = 1.0
    foot_clearance_target: float = 0.04
    foot_height_offset: float = 0.0
    foot_clearance_tracking_sigma: float = 0.01
    about_landing_threshold: float = 0.03
    foot_distance_threshold: float = 0.115  # tron1_pf_config.py:64
    feet_air_time_threshold: float = 0.3   # legged_robot.py:552; 0.25 in go2_ts.py:142
    foot_clearance_uses_terrain: bool = False  # go2_ts.py:147-161 subtracts mean(height_around_feet)
    # link name groups (substring rules, genesis_simulator.py:333-363)
    foot_name: str = "foot"
    penalize_contacts_on: List[str] = field(default_factory=list)
    terminate_after_contacts_on: List[str] = field(default_factory=list)
    obtain_link_contact_states: bool = False
    contact_state_link_names: List[str] = field(default_factory=lambda: ["thigh", "calf", "foot"])
    # periodic-gait framework + sit-pose resets (tron1_pf_ee_config.py:51-63,128-137)
    gait_enabled: bool = False
    gait_period: float = 0.5
    gait_theta_left: float = 0.0
    gait_theta_right: float = 0.5
    gait_b_swing: float = 0.5
    reproduce_r18: bool = False             # bug-compatible env-0 coupling of the periodic-gait tasks; default: per-env semantics
    gait_smooth: bool = False               # gait_function_type "smooth": von Mises CDF indicator (go2_wtw.py:415-453) instead of "step"
    gait_kappa: float = 20.0                # concentration of the von Mises transitions (go2_wtw_config.py:59, tron1_pf_ee_config.py:130)
    base_height_tracking_sigma: float = 0.01
    foot_clearance_mode: int = 0            # height under the foot: 0 none, 1 mean of the 9 samples, 2 max
    sit_init_percent: float = 0.0
    sit_pos: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.55])
    sit_pitch_angle: float = 0.0
    sit_joint_angles: List[float] = field(default_factory=list)
    # go2_wtw behaviour parameters (go2_wtw_config.py:56-73)
    behavior_enabled: bool = False
    behavior_resampling_time: float = 5.0
    gait_period_range: List[float] = field(default_factory=lambda: [0.3, 0.6])
    foot_clearance_target_range: List[float] = field(default_factory=lambda: [0.04, 0.12])
    base_height_target_range: List[float] = field(default_factory=lambda: [0.2, 0.34])
    pitch_target_range: List[float] = field(default_factory=lambda: [-0.3, 0.3])
    gait_theta_lists: List[List[float]] = field(default_factory=list)   # [gait][foot FL,FR,RL,RR]
    euler_tracking_sigma: float = 0.1
    # Constraints as Terminations (go2_cat_config.py:28-36, go2_cat.py:135-215)
    cat_enabled: bool = False
    cat_soft_p: float = 0.25
    cat_action_rate: float = 100.0
    cat_min_base_height: float = 0.25
    cat_max_projected_gravity: float = -0.1
    cat_stand_still_global: bool = True     # as shipped: [N]*[N,1] broadcast couples all envs; False = per env
    double_shift_actions: bool = False      # go2_cat.py:127-130 shifts the action history a second time (R6)
    dof_vel_limits: List[float] = field(default_factory=list)
    # engine knobs (no reference counterpart; DESIGN.md "physics formulation")
    randomize_ctrl_delay: bool = False     # legged_robot.py:240-245: the simulator is driven with a per-env delayed action
    ctrl_delay_step_range: List[int] = field(default_factory=lambda: [0, 1])
    num_teacher: int = 0              # go2_cts: envs [0, num_teacher) are teacher envs (extras only, go2_cts.py:93-99)
    pgs_iterations: int = 12          # sweep cap of the projected block Gauss-Seidel contact solver (truncation error
                                      # at 12 sweeps, mean 6e-5 / p99 1.4e-3 rad/s per substep, sits below the fp32 rounding error of the step)
    pgs_tolerance: float = 1e-4       # stop when max|df| over a sweep <= tol * (1 + max|f|)
    seed: int = 1

    def __post_init__(self):
        if self.foot_clearance_uses_terrain and self.foot_clearance_mode == 0:
            self.foot_clearance_mode = 1

    # ------------------------------------------------------------------ derived
    @property
    def dt(self) -> float:
        return self.sim_dt * self.decimation

    @property
    def num_actions(self) -> int:
        return len(self.dof_names)

    @property
    def max_episode_length(self) -> int:
        return int(math.ceil(self.episode_length_s / self.dt))     # legged_robot.py:446

    @property
    def push_interval(self) -> int:
        return int(math.ceil(self.push_interval_s / self.dt))      # legged_robot.py:455

    @property
    def resample_interval(self) -> int:
        return int(self.resampling_time / self.dt)                 # legged_robot.py:305

    @property
    def fail_limit(self) -> float:
        return self.fail_to_terminal_time_s / self.dt              # legged_robot.py:90

    @property
    def heightfield(self) -> bool:
        return self.mesh_type in ("heightfield", "trimesh")

    @property
    def num_height_points(self) -> int:
        return len(self.measured_points_x) * len(self.measured_points_y)

    def active_rewards(self) -> List[str]:
        """Non-zero terms in evaluation order, 'termination' excluded (legged_robot.py:416-430)."""
        return [n for n in REWARD_TERMS if n != "termination" and self.reward_scales.get(n, 0) != 0]

    def episode_sum_names(self) -> List[str]:
        names = self.active_rewards()
        if self.reward_scales.get("termination", 0) != 0:
            names.append("termination")
        if self.cat_enabled:                       # ConstraintManager.log_all appends these keys (constraint_manager.py:84-91)
            names += ["cstr_" + c for c in CAT_CONSTRAINTS]
        return names

    def scaled_reward(self, name: str) -> np.float32:
        """scale * dt, rounded to fp32 where torch multiplies a python scalar into an fp32 tensor."""
        return np.float32(self.reward_scales.get(name, 0.0) * self.dt)

    def von_mises_series(self):
        """(p, table) of scipy's von Mises CDF series for `gait_kappa` (scipy/stats/_stats.pyx von_mises_cdf, the function
        behind scipy.stats.vonmises.cdf that go2_wtw.py:423-429 calls): p = int(1 + 28 + 0.5 k - 100 / (k + 5)) terms and,
        for n = p-1 .. 1, R_n = 1 / (2 n / k + R_{n+1}) stored as pairs (R_n, R_n / n).  Only kappa < 50 (scipy switches to
        a normal approximation above)."""
        k = float(self.gait_kappa)
        if not 0 < k < 50:
            raise ValueError("the fused 'smooth' gait indicator covers 0 < kappa < 50 (scipy's series branch)")
        p = int(1 + 28.0 + 0.5 * k - 100.0 / (k + 5.0))
        tab, R = [], 0.0
        for n in range(p - 1, 0, -1):
            R = 1.0 / (2.0 * n / k + R)
            tab += [R, R / n]
        return p, np.asarray(tab, np.float64)

    def load_model(self) -> RobotModel:
        return load_robot_model(self.robot, self.dof_names)

    def link_groups(self, model: RobotModel):
        feet = model.find_link_indices([self.foot_name])
        pen = model.find_link_indices(self.penalize_contacts_on)
        term = model.find_link_indices(self.terminate_after_contacts_on)
        cs = model.find_link_indices(self.contact_state_link_names) if self.obtain_link_contact_states else []
        return feet, pen, term, cs

    # observation widths as produced (derived from the actual link count)
    def obs_widths(self, model: RobotModel):
        A = self.num_actions
        feet, _, _, cs = self.link_groups(model)
        if self.obs_kind == "go2":
            return dict(obs=9 + 3 * A, priv=0, single_critic=0, hist=0, critic=0)
        if self.obs_kind == "go2_wtw":                      # go2_wtw.py:53-111
            single = 9 + 3 * A + 2 * len(feet) + 4 + len(feet)
            sc = single + 3 + 7 + 2 * A + len(feet)
            return dict(obs=single, priv=0, single_critic=sc, hist=self.frame_stack * single, critic=self.c_frame_stack * sc)
        if self.obs_kind == "tron1_pf_ee":                  # tron1_pf_ee.py:53-141
            single = 9 + 3 * A + 4
            dr = 10 + 2 * A
            P_ = self.num_height_points if self.measure_heights else 0
            sc = single + dr + 2 + len(cs) + P_ + 3 * len(feet) + 9 * len(feet)
            labels = 3 + len(cs) + len(feet) + 3 * len(feet)
            return dict(obs=single, priv=labels, single_critic=sc, hist=self.frame_stack * single, critic=self.c_frame_stack * sc)
        if self.obs_kind == "tron1_pf":                     # tron1_pf.py:15-70: both outputs are frame stacks
            single = 9 + 3 * A
            sc = 3 + single + A + 7 + len(feet)
            return dict(obs=single, priv=0, single_critic=sc, hist=self.frame_stack * single, critic=self.c_frame_stack * sc)
        if self.obs_kind == "go2_cat":                      # go2_cat.py:19-99: DR info has 3 more entries, no base_lin_vel
            single = 9 + 3 * A
            dr = 10 + 2 * A
            sc = single + dr + len(cs) + (self.num_height_points if self.measure_heights else 0)
            priv = dr + 9 * len(feet) + 3 * len(feet) + len(cs)
            return dict(obs=single, priv=priv, single_critic=sc, hist=self.frame_stack * single, critic=self.c_frame_stack * sc)
        if self.obs_kind in ("go2_ee", "go2_dreamwaq"):    # go2_ee.py:10-73 / go2_dreamwaq.py:7-90: labels instead of priv
            single = 9 + 3 * A
            dr = 7 + 2 * A
            lin = 3 if self.obs_kind == "go2_dreamwaq" else 0
            sc = lin + single + dr + len(cs) + (self.num_height_points if self.measure_heights else 0)
            labels = 3 + len(cs) + len(feet)
            return dict(obs=single, priv=labels, single_critic=sc, hist=self.frame_stack * single, critic=self.c_frame_stack * sc)
        if self.obs_kind in ("go2_ts", "go2_cts"):
            single = 9 + 3 * A
            dr = 7 + 2 * A
            sc = single + dr + 3 + len(cs) + (self.num_height_points if self.measure_heights else 0)
            priv = dr + 9 * len(feet) + 3 * len(feet) + 3 + len(cs)
            return dict(obs=single, priv=priv, single_critic=sc, hist=self.frame_stack * single, critic=self.c_frame_stack * sc)
        raise ValueError(self.obs_kind)

    def noise_scale_vec(self) -> np.ndarray:
        """go2.py:92-117 / go2_ts.py:98-123 (same 45-wide layout)."""
        A = self.num_actions
        if self.obs_kind == "go2_wtw":                     # go2_wtw.py:265-295: zeros beyond the first 9 + 3A entries
            v = np.zeros(9 + 3 * A + 16, np.float32)
            if self.add_noise:
                v[3:6] = self.noise_gravity * self.noise_level
                v[6:9] = self.noise_ang_vel * self.noise_level * self.obs_scale_ang_vel
                v[9:9 + A] = self.noise_dof_pos * self.noise_level * self.obs_scale_dof_pos
                v[9 + A:9 + 2 * A] = self.noise_dof_vel * self.noise_level * self.obs_scale_dof_vel
            return v
        if self.obs_kind == "tron1_pf_ee":
            # tron1_pf_ee.py:309-333 keeps the 12-dof slice bounds on a 31-wide vector: entries 15..20 (dof_vel) get the
            # dof_pos scale and 21..30 (actions, clock) the dof_vel scale -- reproduced as shipped (quirk R17)
            v = np.zeros(9 + 3 * A + 4, np.float32)
            if self.add_noise:
                v[3:6] = self.noise_gravity * self.noise_level
                v[6:9] = self.noise_ang_vel * self.noise_level * self.obs_scale_ang_vel
                v[9:21] = self.noise_dof_pos * self.noise_level * self.obs_scale_dof_pos
                v[21:33] = self.noise_dof_vel * self.noise_level * self.obs_scale_dof_vel
            return v
        v = np.zeros(9 + 3 * A, np.float32)
        if not self.add_noise:
            return v
        v[3:6] = self.noise_gravity * self.noise_level
        v[6:9] = self.noise_ang_vel * self.noise_level * self.obs_scale_ang_vel
        v[9:9 + A] = self.noise_dof_pos * self.noise_level * self.obs_scale_dof_pos
        v[9 + A:9 + 2 * A] = self.noise_dof_vel * self.noise_level * self.obs_scale_dof_vel
        return v

    # ------------------------------------------------------------------ from a live reference cfg
    @classmethod
    def from_reference_cfg(cls, cfg, task: Optional[str] = None) -> "TaskSpec":
        """Read a LeggedGym-Ex config object (nested classes/instances) into a TaskSpec.

        ``task=None`` (plugin mode): the descriptor of the *simulator* -- robot, control, terrain, domain randomisation,
        link groups -- for any task class; reward terms are not read (they stay torch code), the observation layout is
        the plain 9 + 3A one (unused) and the control-delay queue stays with ``LeggedRobot._pre_sim_step``."""
        kinds = {"go2": "go2", "go2_ts": "go2_ts", "go2_cat": "go2_cat", "tron1_pf": "tron1_pf", "tron1_pf_ee": "tron1_pf_ee", "go2_wtw": "go2_wtw",
                 "go2_cts": "go2_cts", "go2_ee": "go2_ee", "go2_dreamwaq": "go2_dreamwaq", PLUGIN_TASK: "go2"}
        plugin = task is None or task == PLUGIN_TASK
        if plugin:
            task = PLUGIN_TASK
        if task not in kinds:
            raise ValueError(f"task {task!r} has no fused descriptor yet (supported: {sorted(kinds)})")

        def to_dict(o):
            return {k: getattr(o, k) for k in dir(o) if not k.startswith("_") and not callable(getattr(o, k))}

        if getattr(getattr(cfg, "sensor", None), "add_depth", False):
            # GenesisSimulator renders depth images when asked to (genesis_simulator.py:135-138, 316-317); this backend has no
            # camera (north_star keeps legged_gym/warp out of the path): refuse instead of handing out a task without its sensor
            raise ValueError("cfg.sensor.add_depth = True: depth sensors are outside the B200 backend")
        t, a, c, d, r, e = cfg.terrain, cfg.asset, cfg.commands, cfg.domain_rand, cfg.rewards, cfg.env
        n, ns = cfg.normalization, cfg.noise
        stiff, damp = cfg.control.stiffness, cfg.control.damping
        kp = {v for k, v in stiff.items() if any(k in dn for dn in a.dof_names)}
        kd = {v for k, v in damp.items() if any(k in dn for dn in a.dof_names)}
        if len(kp) != 1 or len(kd) != 1:
            raise ValueError("per-joint PD gains are not supported by the fused descriptor yet")
        known = [k for k in ("go2", "PF_TRON1A") if k in a.file]
        if not known:
            raise ValueError(f"no packed robot model for asset {a.file!r} (available: go2, PF_TRON1A)")
        robot = {"go2": "go2", "PF_TRON1A": "tron1_pf"}[known[0]]
        hip_thigh_calf = [0.2, 0.4, 0.4] * (len(a.dof_names) // 3)   # go2.py:30-35 / go2_ts.py:86-91
        if task == "tron1_pf":
            hip_thigh_calf = [0.2] * len(a.dof_names)                  # base class, legged_robot.py:279-280
        spec = cls(
            task=task, obs_kind=kinds[task], robot=robot, dof_names=list(a.dof_names),
            num_obs=e.num_observations, num_privileged_obs=e.num_privileged_obs,
            frame_stack=1 if plugin else getattr(e, "frame_stack", 1), c_frame_stack=1 if plugin else getattr(e, "c_frame_stack", 1),
            sim_dt=cfg.sim.dt, decimation=cfg.control.decimation, action_scale=cfg.control.action_scale,
            kp=float(kp.pop()), kd=float(kd.pop()),
            default_dof_pos=[float(cfg.init_state.default_joint_angles[n_]) for n_ in a.dof_names],
            clip_actions=n.clip_actions, clip_observations=n.clip_observations,
            init_pos=list(cfg.init_state.pos), init_quat_xyzw=list(cfg.init_state.rot),
            reset_dof_noise=hip_thigh_calf,
            reset_root_vel=0.0 if task == "go2" else 0.5,
            episode_length_s=e.episode_length_s, fail_to_terminal_time_s=e.fail_to_terminal_time_s,
            max_projected_gravity=r.max_projected_gravity, send_timeouts=e.send_timeouts,
            mesh_type=t.mesh_type, horizontal_scale=t.horizontal_scale, vertical_scale=t.vertical_scale,
            border_size=t.border_size, plane_length=t.plane_length, env_spacing=e.env_spacing,
            terrain_length=t.terrain_length, terrain_width=t.terrain_width, num_rows=t.num_rows, num_cols=t.num_cols,
            terrain_curriculum=bool(t.curriculum) and t.mesh_type in ("heightfield", "trimesh"),
            max_init_terrain_level=t.max_init_terrain_level, static_friction=t.static_friction,
            measure_heights=t.measure_heights, measured_points_x=list(t.measured_points_x),
            measured_points_y=list(t.measured_points_y),
            obtain_terrain_info_around_feet=t.obtain_terrain_info_around_feet,
            num_commands=c.num_commands, resampling_time=c.resampling_time, heading_command=c.heading_command,
            cmd_curriculum=c.curriculum, max_curriculum=c.max_curriculum, curriculum_threshold=c.curriculum_threshold,
            cmd_lin_vel_x=list(c.ranges.lin_vel_x), cmd_lin_vel_y=list(c.ranges.lin_vel_y),
            cmd_ang_vel_yaw=list(c.ranges.ang_vel_yaw), cmd_heading=list(c.ranges.heading),
            randomize_friction=d.randomize_friction, friction_range=list(d.friction_range),
            randomize_base_mass=d.randomize_base_mass, added_mass_range=list(d.added_mass_range),
            push_robots=d.push_robots, push_interval_s=d.push_interval_s, max_push_vel_xy=d.max_push_vel_xy,
            randomize_com_displacement=d.randomize_com_displacement, com_pos_x_range=list(d.com_pos_x_range),
            com_pos_y_range=list(d.com_pos_y_range), com_pos_z_range=list(d.com_pos_z_range),
            randomize_pd_gain=d.randomize_pd_gain, kp_range=list(d.kp_range), kd_range=list(d.kd_range),
            randomize_joint_armature=d.randomize_joint_armature, joint_armature_range=list(d.joint_armature_range),
            randomize_joint_friction=d.randomize_joint_friction, joint_friction_range=list(d.joint_friction_range),
            randomize_joint_damping=d.randomize_joint_damping, joint_damping_range=list(d.joint_damping_range),
            randomize_ctrl_delay=bool(getattr(d, "randomize_ctrl_delay", False)) and not plugin,
            ctrl_delay_step_range=[int(x) for x in getattr(d, "ctrl_delay_step_range", [0, 1])],
            obs_scale_lin_vel=n.obs_scales.lin_vel, obs_scale_ang_vel=n.obs_scales.ang_vel,
            obs_scale_dof_pos=n.obs_scales.dof_pos, obs_scale_dof_vel=n.obs_scales.dof_vel,
            obs_scale_height=n.obs_scales.height_measurements,
            add_noise=ns.add_noise, noise_level=ns.noise_level, noise_dof_pos=ns.noise_scales.dof_pos,
            noise_dof_vel=ns.noise_scales.dof_vel, noise_lin_vel=ns.noise_scales.lin_vel,
            noise_ang_vel=ns.noise_scales.ang_vel, noise_gravity=ns.noise_scales.gravity,
            noise_height=ns.noise_scales.height_measurements,
            reward_scales={} if plugin else {k: float(v) for k, v in to_dict(r.scales).items()},
            only_positive_rewards=r.only_positive_rewards, tracking_sigma=r.tracking_sigma,
            soft_dof_pos_limit=r.soft_dof_pos_limit, base_height_target=r.base_height_target,
            foot_clearance_target=r.foot_clearance_target, foot_height_offset=r.foot_height_offset,
            foot_clearance_tracking_sigma=r.foot_clearance_tracking_sigma,
            about_landing_threshold=getattr(r, "about_landing_threshold", 0.03),
            feet_air_time_threshold=0.25 if task in ("go2_ts", "go2_cat", "tron1_pf", "tron1_pf_ee", "go2_cts", "go2_ee", "go2_dreamwaq") else 0.3,
            foot_distance_threshold=getattr(r, "foot_distance_threshold", 0.115),
            foot_clearance_uses_terrain=task in ("go2_ts", "go2_cat", "go2_cts", "go2_ee", "go2_dreamwaq"),
            dof_vel_limits=list(getattr(a, "dof_vel_limits", []) or []),
            foot_name=a.foot_name, penalize_contacts_on=list(a.penalize_contacts_on),
            terminate_after_contacts_on=list(a.terminate_after_contacts_on),
            obtain_link_contact_states=a.obtain_link_contact_states,
            contact_state_link_names=list(a.contact_state_link_names),
            seed=getattr(cfg, "seed", 1),
        )
        if task == "go2_wtw":
            g, bp = r.periodic_reward_framework, r.behavior_params_range
            spec.gait_smooth, spec.gait_kappa = g.gait_function_type == "smooth", float(g.kappa)
            spec.gait_enabled, spec.behavior_enabled, spec.double_shift_actions = True, True, True
            spec.gait_b_swing = g.b_swing
            spec.gait_theta_lists = [[g.theta_fl_list[k], g.theta_fr_list[k], g.theta_rl_list[k], g.theta_rr_list[k]]
                                     for k in range(len(g.theta_fl_list))]
            spec.behavior_resampling_time = bp.resampling_time
            spec.gait_period_range, spec.foot_clearance_target_range = list(bp.gait_period_range), list(bp.foot_clearance_target_range)
            spec.base_height_target_range, spec.pitch_target_range = list(bp.base_height_target_range), list(bp.pitch_target_range)
            spec.base_height_tracking_sigma, spec.euler_tracking_sigma = r.base_height_tracking_sigma, r.euler_tracking_sigma
            spec.reset_dof_noise = [0.2] * len(a.dof_names)           # base class _reset_dofs, legged_robot.py:279-280
        if task == "tron1_pf_ee":
            g = r.periodic_reward_framework
            spec.gait_smooth, spec.gait_kappa = g.gait_function_type == "smooth", float(g.kappa)
            spec.gait_enabled = True
            spec.gait_period, spec.gait_theta_left, spec.gait_theta_right, spec.gait_b_swing = g.gait_period, g.theta_left, g.theta_right, g.b_swing
            spec.base_height_tracking_sigma = r.base_height_tracking_sigma
            spec.foot_clearance_mode, spec.foot_clearance_uses_terrain = 2, True
            spec.sit_init_percent, spec.sit_pos = cfg.init_state.sit_init_percent, list(cfg.init_state.sit_pos)
            spec.sit_pitch_angle = cfg.init_state.sit_pitch_angle
            spec.sit_joint_angles = [float(cfg.init_state.sit_joint_angles[n_]) for n_ in a.dof_names]
            spec.double_shift_actions = True
            spec.height_obs_offset = 0.6
        if task == "go2_cts":
            spec.foot_clearance_mode = 2                               # go2_cts.py:169: max of the 9 heights under the foot
            spec.num_teacher = int(e.num_teacher)
        if task == "go2_cat":
            cc = cfg.constraints
            spec.cat_enabled = cc.enable == "cat"
            spec.cat_soft_p, spec.cat_action_rate = cc.soft_p, cc.limits.action_rate
            spec.cat_min_base_height, spec.cat_max_projected_gravity = cc.limits.min_base_height, cc.limits.max_projected_gravity
            spec.double_shift_actions = True
        unknown = [k for k, v in spec.reward_scales.items() if v != 0 and k not in REWARD_ID]
        if unknown:
            raise ValueError(f"reward terms without a fused implementation: {unknown}")
        return spec


_GO2_Q0 = [0.0, 0.8, -1.5] * 4


def go2_spec(**over) -> TaskSpec:
    """`go2` flat-terrain velocity tracking (BASELINE config C1; go2_config.py:5-73)."""
    s = TaskSpec(
        task="go2", obs_kind="go2", robot="go2", dof_names=list(GO2_DOF_NAMES), num_obs=45, num_privileged_obs=None,
        default_dof_pos=list(_GO2_Q0), reset_dof_noise=[0.2, 0.4, 0.4] * 4, reset_root_vel=0.0,
        dof_vel_limits=[30.1, 30.1, 15.7] * 4, mesh_type="plane", border_size=5.0, env_spacing=1.0,
        measured_points_x=[round(-0.8 + 0.1 * i, 1) for i in range(17)],
        measured_points_y=[round(-0.5 + 0.1 * i, 1) for i in range(11)],
        cmd_curriculum=True, max_curriculum=1.0, cmd_lin_vel_x=[-0.5, 0.5], cmd_lin_vel_y=[-1.0, 1.0],
        friction_range=[0.5, 1.25], push_interval_s=15.0,
        com_pos_x_range=[-0.01, 0.01], com_pos_y_range=[-0.01, 0.01], com_pos_z_range=[-0.01, 0.01],
        reward_scales=dict(termination=-0.0, tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-0.5,
                           ang_vel_xy=-0.05, orientation=-1.0, torques=-2.0e-4, dof_vel=-5.0e-4, dof_acc=-2.0e-7,
                           base_height=-2.0, feet_air_time=1.0, collision=-1.0, feet_stumble=-0.0, action_rate=-0.01,
                           dof_pos_stand_still=-0.0, dof_pos_limits=-1.0, action_smoothness=-0.01, foot_clearance=0.5),
        soft_dof_pos_limit=0.9, base_height_target=0.36, foot_clearance_target=0.05, foot_height_offset=0.022,
        penalize_contacts_on=["thigh", "calf"], terminate_after_contacts_on=["base", "Head"],
        feet_air_time_threshold=0.3, foot_clearance_uses_terrain=False,
    )
    for k, v in over.items():
        setattr(s, k, v)
    return s


def go2_ts_spec(**over) -> TaskSpec:
    """`go2_ts` rough-terrain teacher-student task (BASELINE config C2; go2_ts_config.py:5-67)."""
    pts = [-0.4, -0.3, -0.2, -0.1, 0.0, 0.1, 0.2, 0.3, 0.4]
    s = TaskSpec(
        task="go2_ts", obs_kind="go2_ts", robot="go2", dof_names=list(GO2_DOF_NAMES), num_obs=45, num_privileged_obs=94,
        frame_stack=20, c_frame_stack=5,
        default_dof_pos=list(_GO2_Q0), reset_dof_noise=[0.2, 0.4, 0.4] * 4, reset_root_vel=0.5,
        dof_vel_limits=[30.1, 30.1, 15.7] * 4, env_spacing=0.5,
        mesh_type="heightfield", border_size=20.0, terrain_length=8.0, terrain_width=8.0, num_rows=10, num_cols=10,
        terrain_curriculum=True, max_init_terrain_level=1, measure_heights=True,
        measured_points_x=list(pts), measured_points_y=list(pts), obtain_terrain_info_around_feet=True,
        cmd_curriculum=True, max_curriculum=1.0, cmd_lin_vel_x=[-0.5, 0.5], cmd_lin_vel_y=[-1.0, 1.0],
        friction_range=[0.2, 1.7], push_interval_s=10.0,
        com_pos_x_range=[-0.03, 0.03], com_pos_y_range=[-0.03, 0.03], com_pos_z_range=[-0.03, 0.03],
        randomize_pd_gain=True, joint_armature_range=[0.015, 0.025], joint_friction_range=[0.01, 0.02],
        joint_damping_range=[0.25, 0.3],
        reward_scales=dict(termination=-0.0, tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-2.0,
                           ang_vel_xy=-0.05, orientation=-0.0, torques=0.0, dof_vel=-0.0, dof_acc=-2.0e-7, base_height=-0.0,
                           feet_air_time=1.0, collision=-1.0, feet_stumble=-0.0, action_rate=-0.01, dof_pos_stand_still=-0.0,
                           dof_pos_limits=-2.0, dof_power=-2.0e-4, action_smoothness=-0.01, foot_clearance=0.2, hip_pos=-0.05,
                           feet_contact_stand_still=0.5),
        soft_dof_pos_limit=0.9, base_height_target=1.0, foot_clearance_target=0.09, foot_height_offset=0.022,
        penalize_contacts_on=["thigh", "calf", "base", "Head", "hip"], terminate_after_contacts_on=[],
        obtain_link_contact_states=True, contact_state_link_names=["thigh", "calf", "foot", "base", "hip"],
        feet_air_time_threshold=0.25, foot_clearance_uses_terrain=True,
    )
    for k, v in over.items():
        setattr(s, k, v)
    return s


def go2_cat_spec(**over) -> TaskSpec:
    """`go2_cat` constraints-as-terminations task (BASELINE config C5; go2_cat_config.py:4-45 on top of Go2TSCfg)."""
    s = go2_ts_spec()
    s.task, s.obs_kind = "go2_cat", "go2_cat"
    s.clip_actions = 10.0
    s.base_height_target = 0.34
    s.reward_scales.update(dof_pos_limits=0.0, collision=0.0, dof_pos_stand_still=0.0, lin_vel_z=-1.0, orientation=-0.5,
                           hip_pos=-0.2, dof_close_to_default=-0.05, foot_clearance=0.2)
    s.cat_enabled, s.double_shift_actions = True, True
    s.dof_vel_limits = [30.1, 30.1, 15.7] * 4
    for k, v in over.items():
        setattr(s, k, v)
    return s


def go2_cts_spec(**over) -> TaskSpec:
    """`go2_cts` concurrent teacher-student (go2_cts_config.py:5-66): go2_ts with raw terrain heights in the privileged
    obs, max-height foot clearance and teacher / student terrain-level logging."""
    s = go2_ts_spec()
    s.task, s.obs_kind = "go2_cts", "go2_cts"
    s.env_spacing = 1.0
    s.joint_friction_range = [0.0, 0.1]
    s.foot_clearance_mode = 2
    s.num_teacher = 4096 // 4 * 3
    for k, v in over.items():
        setattr(s, k, v)
    return s


def go2_ee_spec(**over) -> TaskSpec:
    """`go2_ee` explicit-estimator task (go2_ee_config.py:5-68): estimator features / labels / critic stack."""
    s = go2_ts_spec()
    s.task, s.obs_kind = "go2_ee", "go2_ee"
    s.env_spacing = 2.0
    s.num_obs = 48                          # inherited LeggedRobotEECfg value; the produced frame is 45 wide (obs_widths)
    s.num_privileged_obs = 5 * (45 + 31 + 81 + 17)
    for k, v in over.items():
        setattr(s, k, v)
    return s


def go2_dreamwaq_spec(**over) -> TaskSpec:
    """`go2_dreamwaq` (go2_dreamwaq_config.py:5-70): obs, critic stack, history, explicit labels, next state."""
    s = go2_ts_spec()
    s.task, s.obs_kind = "go2_dreamwaq", "go2_dreamwaq"
    s.env_spacing = 1.0
    s.num_privileged_obs = 5 * (45 + 31 + 81 + 17 + 3)
    for k, v in over.items():
        setattr(s, k, v)
    return s


def tron1_pf_spec(**over) -> TaskSpec:
    """`tron1_pf` point-foot biped, flat terrain (tron1_pf_config.py:4-117): 2 chains x 3 joints, frame-stacked obs."""
    s = TaskSpec(
        task="tron1_pf", obs_kind="tron1_pf", robot="tron1_pf", dof_names=list(TRON1_PF_DOF_NAMES), num_obs=135,
        num_privileged_obs=225, frame_stack=5, c_frame_stack=5, kp=42.0, kd=2.5, action_scale=0.25,
        default_dof_pos=[0.0] * 6, reset_dof_noise=[0.2] * 6, reset_root_vel=0.5, init_pos=[0.0, 0.0, 0.8],
        mesh_type="plane", border_size=5.0, env_spacing=2.0,
        measured_points_x=[round(-0.8 + 0.1 * i, 1) for i in range(17)],
        measured_points_y=[round(-0.5 + 0.1 * i, 1) for i in range(11)],
        cmd_curriculum=True, max_curriculum=1.0, cmd_lin_vel_x=[-0.5, 0.5], cmd_lin_vel_y=[-0.6, 0.6],
        friction_range=[0.5, 1.25], push_interval_s=10.0,
        com_pos_x_range=[-0.03, 0.03], com_pos_y_range=[-0.03, 0.03], com_pos_z_range=[-0.03, 0.03],
        reward_scales=dict(termination=-0.0, tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-0.5, ang_vel_xy=-0.05,
                           orientation=-3.0, torques=-2.0e-5, dof_vel=-5.0e-4, dof_acc=-2.0e-7, base_height=-2.0,
                           feet_air_time=1.0, collision=-1.0, feet_stumble=-0.0, action_rate=-0.01, dof_pos_stand_still=-0.0,
                           keep_balance=1.0, dof_pos_limits=-2.0, feet_distance=-100.0, action_smoothness=-0.01,
                           foot_clearance=0.5, no_fly=0.5, foot_landing_vel=-0.15),
        only_positive_rewards=False, soft_dof_pos_limit=0.9, base_height_target=0.68, foot_clearance_target=0.07,
        foot_height_offset=0.032, about_landing_threshold=0.1, foot_distance_threshold=0.115,
        penalize_contacts_on=["knee", "hip"], terminate_after_contacts_on=["base", "abad"],
        feet_air_time_threshold=0.25, foot_clearance_uses_terrain=False,
    )
    for k, v in over.items():
        setattr(s, k, v)
    return s


def tron1_pf_ee_spec(**over) -> TaskSpec:
    """`tron1_pf_ee` point-foot biped on rough terrain with full DR, periodic gait and sit-pose resets (BASELINE config
    C4; tron1_pf_ee_config.py:4-170)."""
    pts = [-0.3, -0.2, -0.1, 0.0, 0.1, 0.2, 0.3]
    s = TaskSpec(
        task="tron1_pf_ee", obs_kind="tron1_pf_ee", robot="tron1_pf", dof_names=list(TRON1_PF_DOF_NAMES), num_obs=48,
        num_privileged_obs=1340, frame_stack=10, c_frame_stack=10, kp=42.0, kd=2.5, action_scale=0.25, clip_actions=20.0,
        default_dof_pos=[0.0] * 6, reset_dof_noise=[0.2, 0.4, 0.4] * 2, reset_root_vel=0.5, init_pos=[0.0, 0.0, 0.83],
        env_spacing=3.0, mesh_type="heightfield", border_size=15.0, terrain_length=8.0, terrain_width=8.0, num_rows=10,
        num_cols=10, terrain_curriculum=True, max_init_terrain_level=1, measure_heights=True,
        measured_points_x=list(pts), measured_points_y=list(pts), obtain_terrain_info_around_feet=True,
        cmd_curriculum=True, max_curriculum=0.8, cmd_lin_vel_x=[-0.5, 0.5], cmd_lin_vel_y=[-0.6, 0.6],
        friction_range=[0.0, 1.7], added_mass_range=[-1.0, 2.0], push_interval_s=10.0,
        com_pos_x_range=[-0.03, 0.03], com_pos_y_range=[-0.03, 0.03], com_pos_z_range=[-0.03, 0.03],
        randomize_pd_gain=True, randomize_joint_armature=True, joint_armature_range=[0.11, 0.13],
        randomize_joint_friction=True, joint_friction_range=[0.0, 0.01], randomize_joint_damping=True,
        joint_damping_range=[1.4, 1.45],
        reward_scales=dict(termination=-0.0, tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-0.5, ang_vel_xy=-0.05,
                           orientation=-4.0, torques=0.0, dof_vel=-0.0, dof_acc=-2.0e-7, base_height=-0.0, feet_air_time=0.0,
                           collision=-1.0, feet_stumble=-0.0, action_rate=-0.01, dof_pos_stand_still=-0.0, keep_balance=1.0,
                           dof_pos_limits=-2.0, feet_distance=-100.0, tracking_base_height=0.3, dof_power=-2.0e-4,
                           foot_acc=-1.0e-5, action_smoothness=-0.01, biped_periodic_gait=1.0, foot_clearance=0.5),
        only_positive_rewards=False, soft_dof_pos_limit=0.95, base_height_target=0.75, foot_clearance_target=0.06,
        foot_height_offset=0.032, foot_distance_threshold=0.115, max_projected_gravity=-0.2,
        penalize_contacts_on=["knee", "hip"], terminate_after_contacts_on=["base", "abad"],
        obtain_link_contact_states=True, contact_state_link_names=["hip", "knee", "foot"],
        feet_air_time_threshold=0.25, foot_clearance_uses_terrain=True, foot_clearance_mode=2, height_obs_offset=0.6,
        gait_enabled=True, gait_period=0.5, gait_theta_left=0.0, gait_theta_right=0.5, gait_b_swing=0.5,
        base_height_tracking_sigma=0.01, sit_init_percent=0.7, sit_pos=[0.0, 0.0, 0.55], sit_pitch_angle=-0.2,
        sit_joint_angles=[0.0, 0.6, 1.36, 0.0, -0.6, -1.36], double_shift_actions=True,
    )
    for k, v in over.items():
        setattr(s, k, v)
    return s


def go2_wtw_spec(**over) -> TaskSpec:
    """`go2_wtw` Walk-These-Ways on flat ground with periodic-gait rewards and behaviour parameters (BASELINE config C3;
    go2_wtw_config.py:5-103)."""
    s = TaskSpec(
        task="go2_wtw", obs_kind="go2_wtw", robot="go2", dof_names=list(GO2_DOF_NAMES), num_obs=305, num_privileged_obs=495,
        frame_stack=5, c_frame_stack=5, default_dof_pos=list(_GO2_Q0), reset_dof_noise=[0.2] * 12, reset_root_vel=0.5,
        dof_vel_limits=[30.1, 30.1, 15.7] * 4, mesh_type="plane", border_size=5.0, env_spacing=1.0,
        measured_points_x=[round(-0.8 + 0.1 * i, 1) for i in range(17)],
        measured_points_y=[round(-0.5 + 0.1 * i, 1) for i in range(11)],
        cmd_curriculum=True, max_curriculum=1.0, resampling_time=8.0, cmd_lin_vel_x=[-0.5, 0.5], cmd_lin_vel_y=[-1.0, 1.0],
        friction_range=[0.2, 1.7], push_interval_s=15.0, randomize_pd_gain=True,
        com_pos_x_range=[-0.03, 0.03], com_pos_y_range=[-0.03, 0.03], com_pos_z_range=[-0.03, 0.03],
        reward_scales=dict(termination=-0.0, tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-0.5, ang_vel_xy=-0.05,
                           orientation=-0.0, torques=-2.0e-4, dof_vel=-5.0e-4, dof_acc=-2.0e-7, base_height=-0.0, feet_air_time=0.0,
                           collision=-1.0, feet_stumble=-0.0, action_rate=-0.01, dof_pos_stand_still=-0.0, dof_pos_limits=-10.0,
                           tracking_base_height=0.6, tracking_orientation=0.6, tracking_foot_clearance=0.9,
                           quad_periodic_gait=1.5, action_smoothness=-0.01, foot_landing_vel=-0.1, hip_pos=-1.0),
        only_positive_rewards=True, soft_dof_pos_limit=0.9, foot_height_offset=0.022, about_landing_threshold=0.03,
        base_height_tracking_sigma=0.01, euler_tracking_sigma=0.1,
        penalize_contacts_on=["thigh", "calf"], terminate_after_contacts_on=["base", "Head"],
        gait_enabled=True, behavior_enabled=True, double_shift_actions=True, gait_b_swing=0.5,
        gait_theta_lists=[[0.0, 0.5, 0.5, 0.0], [0.0, 0.0, 0.0, 0.0], [0.5, 0.0, 0.5, 0.0], [0.0, 0.0, 0.5, 0.5]],
        behavior_resampling_time=5.0,
    )
    for k, v in over.items():
        setattr(s, k, v)
    return s


PRESETS = {"go2": go2_spec, "go2_ts": go2_ts_spec, "go2_cat": go2_cat_spec, "tron1_pf": tron1_pf_spec,
           "tron1_pf_ee": tron1_pf_ee_spec, "go2_wtw": go2_wtw_spec, "go2_cts": go2_cts_spec, "go2_ee": go2_ee_spec,
           "go2_dreamwaq": go2_dreamwaq_spec}
