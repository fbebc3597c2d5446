/*
 * b200_step.h -- C ABI of the B200-native "physics + environment step" backend.
 *
 * One shared library (hcr_genesis_lr_cl_b200/libb200step.so), plain pointers and sizes, no torch types.
 * Every entry point replaces a piece of the reference's Python hot path (LeggedGym-Ex, paths relative
 * to the reference root):
 *
 *   b200_dynamics_step   LeggedRobot._pre_sim_step            legged_gym/envs/base/legged_robot.py:230-252
 *                        GenesisSimulator.step                legged_gym/simulator/genesis_simulator.py:20-33
 *                        GenesisSimulator._compute_torques    legged_gym/simulator/genesis_simulator.py:630-642
 *                        scene.step() (third-party engine)    legged_gym/simulator/genesis_simulator.py:29
 *   b200_simulator_step  GenesisSimulator.step alone          legged_gym/simulator/genesis_simulator.py:20-33 (plugin mode:
 *                        the stock LeggedRobot._pre_sim_step has already clipped / delayed the actions it passes in)
 *   b200_env_post_step   LeggedRobot.post_physics_step        legged_gym/envs/base/legged_robot.py:55-76
 *                        GenesisSimulator.post_physics_step   legged_gym/simulator/genesis_simulator.py:35-60
 *                        (+ everything those call: height scan :552-610, OOB :612-628, callbacks,
 *                        check_termination, compute_reward, reset_idx, compute_observations)
 *   b200_env_step        LeggedRobot.step, all of the above   legged_gym/envs/base/legged_robot.py:37-53, with HOST buffers:
 *                        + the host copies of the rollout     rsl_rl/runners/on_policy_runner.py:118-139
 *   b200_reset_all       BaseTask.reset -> reset_idx(all)     legged_gym/envs/base/base_task.py:60-64
 *
 * All state lives in caller-owned device buffers (torch CUDA tensors in the Python host layer) that are
 * bound once with b200_bind_buffers(); the Simulator plugin properties (simulator.py:243-613) are
 * zero-copy views of the same memory.  Functions return 0 on success, non-zero on failure with a
 * message in b200_last_error() (the Python wrapper raises RuntimeError).  No host synchronisation
 * happens inside the step functions; they enqueue on the stream they are given.
 *
 * The enums below index the flat task descriptor (float[] / int[]); hcr_genesis_lr_cl_b200/_cabi.py parses
 * this header, so the names here are the single source of truth for the host layer too.
 */
#ifndef B200_STEP_H
#define B200_STEP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_MAX_JOINTS 12
#define B200_MAX_CHAINS 4
#define B200_MAX_BODIES 13
#define B200_MAX_LINKS 17
#define B200_MAX_SPHERES 64
#define B200_MAX_FEET 4
#define B200_KMAX 8        /* contacts kept per env                        */
#define B200_AUXMAX 8      /* joint-limit + frictionloss constraint rows   */
#define B200_RMAX 32       /* constraint rows = 3*KMAX + AUXMAX = one warp */
#define B200_MAX_REWARDS 33
#define B200_MAX_GAITS 4
#define B200_STATS_EXTRA 4    /* per-step statistics beyond the episode sums: terrain level, cstr prob | teacher level, student level, envs reset */
#define B200_MAX_CTRL_DELAY 7 /* largest ctrl_delay_step_range[1] */
#define B200_VM_MAX 56       /* most terms of the von Mises CDF series (kappa < 50, "smooth" gait indicator) */
#define B200_GAIT_STATE 20   /* floats per env in gait_state */
/* columns of one gait_state row: theta[4] (FL,FR,RL,RR or L,R), gait time, phase, gait period, base-height /
 * foot-clearance / pitch targets (go2_wtw behaviour parameters), clock[8] = sin[feet], cos[feet] */
#define B200_GS_TH 0
#define B200_GS_GT 4
#define B200_GS_PHI 5
#define B200_GS_PER 6
#define B200_GS_BH 7
#define B200_GS_FC 8
#define B200_GS_PT 9
#define B200_GS_CLK 10
#define B200_MAX_OBS 64
#define B200_MAX_PTS_AXIS 17
#define B200_BODY_STRIDE 20

/* ---- task descriptor: float section ---- */
enum B200TaskF {
    TF_SIM_DT, TF_POLICY_DT, TF_ACTION_SCALE, TF_KP, TF_KD, TF_CLIP_ACTIONS, TF_CLIP_OBS,
    TF_INIT_POS, TF_INIT_POS_1, TF_INIT_POS_2, TF_INIT_QUAT, TF_INIT_QUAT_1, TF_INIT_QUAT_2, TF_INIT_QUAT_3, /* xyzw */
    TF_RESET_ROOT_XY, TF_RESET_ROOT_VEL, TF_FAIL_LIMIT, TF_MAX_PROJ_GRAV,
    TF_HSCALE, TF_VSCALE, TF_BORDER, TF_X_LO, TF_X_HI, TF_Y_LO, TF_Y_HI, TF_TERRAIN_HALF_LENGTH, TF_EPISODE_LENGTH_S,
    /* uniform ranges are stored as (lower, span) with span = fp32(upper - lower) evaluated in double, which is
       what torch_rand_float does with its python-float arguments (math_utils.py:78-81) */
    TF_CMD_VY_LO, TF_CMD_VY_SPAN, TF_CMD_YAW_LO, TF_CMD_YAW_HI, TF_CMD_YAW_SPAN, TF_CMD_HEADING_LO, TF_CMD_HEADING_SPAN,
    TF_FRICTION_LO, TF_FRICTION_SPAN, TF_MASS_LO, TF_MASS_SPAN, TF_COMX_LO, TF_COMX_SPAN, TF_COMY_LO, TF_COMY_SPAN,
    TF_COMZ_LO, TF_COMZ_SPAN, TF_KPS_LO, TF_KPS_SPAN, TF_KDS_LO, TF_KDS_SPAN, TF_ARM_LO, TF_ARM_SPAN, TF_JFR_LO, TF_JFR_SPAN,
    TF_JDA_LO, TF_JDA_SPAN, TF_MAX_PUSH, TF_FRICTION_OFFSET, TF_KPS_OFFSET, TF_KDS_OFFSET,
    TF_OS_LIN_VEL, TF_OS_ANG_VEL, TF_OS_DOF_POS, TF_OS_DOF_VEL, TF_OS_HEIGHT, TF_HEIGHT_OBS_OFFSET,
    TF_TRACKING_SIGMA, TF_BASE_HEIGHT_TARGET, TF_FOOT_CLEARANCE_TARGET, TF_FOOT_HEIGHT_OFFSET, TF_FOOT_CLEARANCE_SIGMA,
    TF_ABOUT_LANDING, TF_AIR_TIME_THRESHOLD, TF_FOOT_DISTANCE_THRESHOLD,
    TF_GRAV, TF_TC, TF_DAMPRATIO, TF_D0, TF_DMAX, TF_WIDTH, TF_MID, TF_POWER, TF_TERRAIN_MU, TF_GEOM_MU, TF_PGS_TOL,
    /* Constraints-as-Terminations (go2_cat.py:135-215, go2_cat_config.py:28-36) */
    TF_CAT_SOFT_P, TF_CAT_ACTION_RATE, TF_CAT_MIN_BASE_HEIGHT, TF_CAT_MAX_PROJ_GRAV,
    /* periodic-gait framework and sit-pose resets (tron1_pf_ee.py:186-256,347-437; tron1_pf_ee_config.py:51-63,128-137) */
    TF_GAIT_PERIOD, TF_GAIT_THETA_LEFT, TF_GAIT_THETA_RIGHT, TF_GAIT_B_SWING, TF_BASE_HEIGHT_SIGMA, TF_SIT_PERCENT,
    TF_SIT_POS, TF_SIT_POS_1, TF_SIT_POS_2, TF_SIT_QUAT, TF_SIT_QUAT_1, TF_SIT_QUAT_2, TF_SIT_QUAT_3, /* xyzw */
    TF_EULER_SIGMA,                                             /* go2_wtw tracking_orientation */
    TF_DEFAULT_DOF_POS,                                         /* [B200_MAX_JOINTS] */
    TF_RESET_DOF_NOISE = TF_DEFAULT_DOF_POS + B200_MAX_JOINTS,  /* [B200_MAX_JOINTS] */
    TF_DOF_LIM_LO = TF_RESET_DOF_NOISE + B200_MAX_JOINTS,       /* [B200_MAX_JOINTS] soft limits */
    TF_DOF_LIM_HI = TF_DOF_LIM_LO + B200_MAX_JOINTS,            /* [B200_MAX_JOINTS] */
    TF_REWARD_SCALE = TF_DOF_LIM_HI + B200_MAX_JOINTS,          /* [B200_MAX_REWARDS] scale*dt, by term id */
    TF_NOISE_VEC = TF_REWARD_SCALE + B200_MAX_REWARDS,          /* [B200_MAX_OBS] */
    TF_POINTS_X = TF_NOISE_VEC + B200_MAX_OBS,                  /* [B200_MAX_PTS_AXIS] */
    TF_POINTS_Y = TF_POINTS_X + B200_MAX_PTS_AXIS,              /* [B200_MAX_PTS_AXIS] */
    TF_TORQUE_LIMIT = TF_POINTS_Y + B200_MAX_PTS_AXIS,          /* [B200_MAX_JOINTS] URDF effort (simulator.torque_limits) */
    TF_DOF_VEL_LIMIT = TF_TORQUE_LIMIT + B200_MAX_JOINTS,       /* [B200_MAX_JOINTS] cfg.asset.dof_vel_limits */
    TF_SIT_DOF_POS = TF_DOF_VEL_LIMIT + B200_MAX_JOINTS,        /* [B200_MAX_JOINTS] sit-pose joint angles */
    TF_GAIT_THETA = TF_SIT_DOF_POS + B200_MAX_JOINTS,           /* [B200_MAX_GAITS][B200_MAX_FEET] phase offsets per gait (go2_wtw) */
    /* "smooth" periodic-gait indicator (go2_wtw.py:415-453, tron1_pf_ee.py:369-407): the von Mises CDF that the reference
       gets from scipy.stats.vonmises.cdf on the host, as scipy's own series -- per term n = p-1 .. 1 the pair
       (R_n, R_n / n) of the backward recursion R_n = 1 / (2 n / kappa + R_{n+1}), which depends on kappa alone */
    TF_VM_R = TF_GAIT_THETA + B200_MAX_GAITS * B200_MAX_FEET,   /* [2 * B200_VM_MAX] */
    TF_COUNT = TF_VM_R + 2 * B200_VM_MAX
};

/* ---- task descriptor: int section ---- */
enum B200TaskI {
    TI_NUM_ENVS, TI_A, TI_C, TI_D, TI_L, TI_F, TI_PX, TI_PY, TI_NSPHERES, TI_DECIMATION, TI_PGS_ITERS,
    TI_OBS_KIND, TI_NUM_OBS, TI_NUM_PRIV, TI_SINGLE_CRITIC, TI_FRAME_STACK, TI_C_FRAME_STACK,
    TI_MAX_EPISODE_LENGTH, TI_RESAMPLE_INTERVAL, TI_PUSH_INTERVAL,
    TI_HEIGHTFIELD, TI_MEASURE_HEIGHTS, TI_FEET_INFO, TI_TERRAIN_CURRICULUM, TI_HEADING_COMMAND, TI_PUSH_ROBOTS,
    TI_ADD_NOISE, TI_ONLY_POSITIVE, TI_RAND_FRICTION, TI_RAND_MASS, TI_RAND_COM, TI_RAND_PD, TI_RAND_ARMATURE,
    TI_RAND_JFRICTION, TI_RAND_JDAMPING, TI_CONTACT_STATES, TI_CLEARANCE_USES_TERRAIN,
    TI_NUM_LEVELS, TI_NUM_TYPES, TI_HF_ROWS, TI_HF_COLS, TI_SEED_LO, TI_SEED_HI,
    TI_N_REWARDS, TI_TERMINATION_COL,                            /* column of "termination" in episode_sums or -1 */
    TI_N_PEN, TI_N_TERM, TI_N_CS,
    TI_ENV_OFFSET,                                               /* global id of local env 0 (multi-GPU sharding; keys the RNG) */
    TI_CAT,                                                      /* 1: compute CaT constraint probabilities (9 constraints) */
    TI_CAT_GLOBAL_STANDSTILL,                                    /* 1: reproduce the [N,N] broadcast of go2_cat.py:177-178 */
    TI_DOUBLE_SHIFT,                                             /* 1: tasks that shift the action history again after the step (R6) */
    TI_GAIT,                                                     /* 1: biped periodic-gait state (theta, gait time, phi, clock) */
    TI_BEHAVIOR,                                                 /* 1: per-env behaviour params (gait period, targets) resampled (go2_wtw) */
    TI_BEHAVIOR_INTERVAL,                                        /* int(behavior resampling_time / dt) */
    TI_CTRL_DELAY,                                               /* 1: per-env action delay queue (legged_robot.py:240-245) */
    TI_CTRL_DELAY_LO, TI_CTRL_DELAY_HI,                          /* domain_rand.ctrl_delay_step_range (inclusive); queue depth = HI + 1 */
    TI_NUM_TEACHER,                                              /* go2_cts: global envs [0, num_teacher) are teacher envs (go2_cts.py:93-99) */
    TI_CLEARANCE_MODE,                                           /* foot clearance / labels relative to: 0 nothing, 1 mean, 2 max of the 9 heights */
    TI_N_SUMS,                                                   /* columns of episode_sums: rewards (+termination) (+9 cstr_*) */
    TI_R18,                                                      /* 1: reproduce the reference's env-0 coupling of the periodic-gait tasks: flattened
                                                                    nonzero() of [N,1] masks makes env 0 follow "any env" for the gait-clock wrap and the swing /
                                                                    stance indicator (go2_wtw.py:258-263,443-466, tron1_pf_ee.py:27-35,397-420); default 0 = per env */
    TI_VM_TERMS,                                                 /* 0: "step" gait indicator; p > 0: "smooth" indicator, von Mises CDF series of p terms */
    TI_TRIMESH,                                                  /* terrain.mesh_type "trimesh": collide with the triangles convert_heightfield_to_trimesh builds
                                                                    (legged_gym/utils/terrain_utils.py:887-900: cell diagonal (i,j)-(i+1,j+1)); 0 = heightfield */
    TI_REWARD_IDS,                                               /* [B200_MAX_REWARDS] active term ids, evaluation order */
    TI_FEET_LINKS = TI_REWARD_IDS + B200_MAX_REWARDS,            /* [B200_MAX_FEET]  */
    TI_PEN_LINKS = TI_FEET_LINKS + B200_MAX_FEET,                /* [B200_MAX_LINKS] */
    TI_TERM_LINKS = TI_PEN_LINKS + B200_MAX_LINKS,               /* [B200_MAX_LINKS] */
    TI_CS_LINKS = TI_TERM_LINKS + B200_MAX_LINKS,                /* [B200_MAX_LINKS] */
    TI_COUNT = TI_CS_LINKS + B200_MAX_LINKS
};

/* reward term ids (alphabetical == the reference's evaluation order, legged_robot.py:411-434) */
enum B200Reward {
    RW_ACTION_RATE, RW_ACTION_SMOOTHNESS, RW_ANG_VEL_XY, RW_BASE_HEIGHT, RW_BIPED_PERIODIC_GAIT, RW_COLLISION, RW_DOF_ACC,
    RW_DOF_CLOSE_TO_DEFAULT, RW_DOF_POS_LIMITS, RW_DOF_POS_STAND_STILL, RW_DOF_POWER, RW_DOF_VEL,
    RW_DOF_VEL_STAND_STILL, RW_FEET_AIR_TIME, RW_FEET_CONTACT_STAND_STILL, RW_FEET_DISTANCE, RW_FOOT_ACC, RW_FOOT_CLEARANCE,
    RW_FOOT_LANDING_VEL, RW_HIP_POS, RW_KEEP_BALANCE, RW_LIN_VEL_Z, RW_NO_FLY, RW_ORIENTATION, RW_QUAD_PERIODIC_GAIT, RW_THIGH_POS, RW_TORQUES,
    RW_TRACKING_ANG_VEL, RW_TRACKING_BASE_HEIGHT, RW_TRACKING_FOOT_CLEARANCE, RW_TRACKING_LIN_VEL, RW_TRACKING_ORIENTATION,
    RW_TERMINATION, RW_COUNT
};

/* random-draw sites: u = philox4x32-10(key = seed, counter = (env, step, site, idx >> 2))[idx & 3] */
enum B200Site {
    SITE_CMD_RESAMPLE, SITE_PUSH, SITE_LEVEL, SITE_CMD_RESET, SITE_DOF, SITE_ROOT, SITE_FRICTION, SITE_MASS, SITE_COM,
    SITE_KP, SITE_KD, SITE_ARMATURE, SITE_JFRICTION, SITE_JDAMPING, SITE_OBS_NOISE,
    SITE_GAIT,        /* idx 0: theta offset, 1: gait time (tron1_pf_ee.py:222-226) */
    SITE_HOST,        /* host-side scalar draws keyed env = 0xffffffff: idx 0 = sit-pose coin (R8), 1 / 2 = gait choice of the
                         callback / reset resampling of the step (R7) */
    SITE_BEHAVIOR,    /* idx 0..3: gait period, base height, foot clearance, pitch targets (callback resampling) */
    SITE_BEHAVIOR_RESET,
    SITE_CTRL_DELAY   /* idx 0: action delay of an env that resets (legged_robot.py:144-148) */
};

/* phases of b200_env_post_step (bit mask) */
enum B200Phase {
    PHASE_SIM_POST = 1,      /* GenesisSimulator.post_physics_step: derived base state, contacts, height scan */
    PHASE_CALLBACK = 2,      /* command resampling, heading command, pushes */
    PHASE_TERMINATION = 4,
    PHASE_REWARD = 8,
    PHASE_RESET = 16,
    PHASE_OBSERVE = 32,
    PHASE_ALL = 63
};

/*
 * Device buffers, row-major [num_envs, k] float32 unless noted.  The field order below is the ABI
 * (the host layer fills a ctypes array of pointers in this order).
 */
typedef struct B200Buffers {
    /* physics state */
    float *base_pos;            /* [N,3]  world; also the API `base_pos`                              */
    float *base_quat_wxyz;      /* [N,4]  engine convention (genesis_simulator.py:40)                 */
    float *base_lin_w;          /* [N,3]  world-frame linear velocity (dofs 0..2)                     */
    float *base_ang_w;          /* [N,3]  world-frame angular velocity (dofs 3..5)                    */
    float *dof_pos;             /* [N,A]  joint positions, cfg.asset.dof_names order                  */
    float *dof_vel;             /* [N,A]                                                              */
    /* domain randomisation (private attrs the tasks read) */
    float *friction;            /* [N,1]  _friction_values     */
    float *added_mass;          /* [N,1]  _added_base_mass     */
    float *com_bias;            /* [N,3]  _base_com_bias       */
    float *kp_scale;            /* [N,A]  _kp_scale            */
    float *kd_scale;            /* [N,A]  _kd_scale            */
    float *joint_armature;      /* [N,1]  _joint_armature      */
    float *joint_friction;      /* [N,1]  _joint_friction      */
    float *joint_damping;       /* [N,1]  _joint_damping       */
    float *rand_push_vels;      /* [N,3]  _rand_push_vels      */
    /* task state */
    float *actions;             /* [N,A] */
    float *last_actions;        /* [N,A] */
    float *llast_actions;       /* [N,A] */
    float *commands;            /* [N,4] */
    int32_t *episode_length;    /* [N] int32 (episode_length_buf) */
    int32_t *fail_buf;          /* [N] int32 */
    float *feet_air_time;       /* [N,F] */
    uint8_t *last_contacts;     /* [N,F] uint8 */
    float *episode_sums;        /* [N,n_sums] columns = active reward terms (+ termination) */
    int64_t *terrain_levels;    /* [N] int64 */
    int64_t *terrain_types;     /* [N] int64 */
    float *env_origins;         /* [N,3] */
    /* derived, API visible */
    float *base_quat;           /* [N,4] xyzw */
    float *base_euler;          /* [N,3] */
    float *base_lin_vel;        /* [N,3] body frame */
    float *base_ang_vel;        /* [N,3] body frame */
    float *projected_gravity;   /* [N,3] */
    float *torques;             /* [N,A] unclipped PD torque of the last substep */
    float *link_contact_forces; /* [N,L,3] */
    float *feet_pos;            /* [N,F,3] */
    float *feet_vel;            /* [N,F,3] */
    float *link_contact_states; /* [N,n_cs] */
    float *measured_heights;    /* [N,P] */
    float *height_around_feet;  /* [N,F,9] */
    float *normal_vector_around_feet; /* [N,3F] */
    float *last_dof_vel;        /* [N,A] */
    float *last_feet_vel;       /* [N,F,3] */
    float *last_base_lin_vel;   /* [N,3] */
    float *last_base_ang_vel;   /* [N,3] */
    /* outputs */
    float *obs_buf;             /* [N,num_obs]   */
    float *privileged_obs_buf;  /* [N,num_priv]  */
    float *obs_history;         /* [N, 2*(frame_stack+1), num_obs]   frame stack as a double-written ring of period M = K + 1: the frame of
                                   observation step t sits in slots t mod M and t mod M + M, so the K most recent frames (oldest first: the
                                   tensor the reference concatenates from its deque, legged_robot_ts.py:29-47) are the contiguous slots
                                   [t mod M + 2, t mod M + M]; the spare slot keeps the window of step t intact while step t + 1 writes */
    float *critic_obs;          /* [N, 2*(c_frame_stack+1), single_critic]   same, for the critic stack */
    float *rew_buf;             /* [N] */
    uint8_t *reset_buf;         /* [N] uint8 (torch.bool storage) */
    uint8_t *time_out_buf;      /* [N] uint8 */
    int32_t *height_cells;      /* [N,P,2] int32 cell indices of the scan (diagnostic, may be NULL) */
    float *stats;               /* [2*n_sums+4 + 32*(n_sums+B200_STATS_EXTRA)] per-step reductions: [0,n_sums) sum over resetting envs of
                                   episode_sums, [n_sums] count of resets, [n_sums+1] sum of terrain levels (all envs);
                                   [n_sums+2] sum of cstr_prob (all envs), [n_sums+3] sum of the teacher envs' terrain levels; then a
                                   ring of 32 slots [n_sums+B200_STATS_EXTRA]: the extras["episode"] means of step % 32 (rew_* in
                                   episode-sum order, mean terrain level, mean cstr_prob | teacher level, student level, envs reset) */
    float *gait_state;          /* [N,20] theta[4], gait_time, phi, gait_period, base_height_target, foot_clearance_target,
                                   pitch_target, clock_input[8] = sin[F], cos[F] (periodic-gait tasks) */
    float *cstr_prob;           /* [N] CaT termination probability (Go2CaT.cstr_prob); unused (but bound) for other tasks */
    int32_t *global_flags;      /* [4] int32: [0] bit 0: any env with |dof_vel| > 4 after the physics step (CaT stand-still, R4), bits 8..16
                                   (TI_R18): any env wraps its gait clock this step | any env in swing, per foot | in stance, per foot; [1] ticket of the
                                   env kernel's last-CTA finalize; [2] envs reset because their state went non-finite (cumulative) */
    float *contact_warm;        /* [N,48] contact-solver warm start carried between substeps and policy steps: 8 x (sphere id + 1,
                                   f_n, f_t1, f_t2) then 8 x (aux-row code + 1, f); zero = empty */
    float *next_state_buf;      /* [N,num_obs] go2_dreamwaq decoder target (go2_dreamwaq.py:72-80); [N,1] otherwise */
    float *action_queue;        /* [N, ctrl_delay_hi + 1, A] clipped actions of the last steps, newest first (legged_robot.py:240-245); [N,1] when off */
    int32_t *action_delay;      /* [N] slot of action_queue the simulator is driven with (re-drawn on reset, legged_robot.py:144-148) */
    /* scheduling state of the dynamics kernel (no reference counterpart; results do not depend on it): per-env cost of the
     * last two launches and the env each warp slot takes in the next ones */
    int32_t *dyn_cost;          /* [2,N] solver work (sweeps x rows) of each env in the launches of either parity */
    int32_t *dyn_order;         /* [2,N] warp slot w runs env w + dyn_order[parity][w] (delta-encoded permutation: zeros = identity) */
    /* failure containment (no reference counterpart: the reference never checks): an env whose state left
     * the finite range inside the dynamics kernel keeps its last finite pose with zero velocities, is flagged here, and the
     * next b200_env_post_step that runs PHASE_TERMINATION resets it (counted in global_flags[2]); in plugin mode the flag is
     * exposed as B200Simulator.nonfinite_envs for the task to act on */
    int32_t *nonfinite;         /* [N] int32 */
} B200Buffers;

typedef struct B200Handle B200Handle;

int b200_create(const int32_t *model_i, int n_model_i, const float *model_f, int n_model_f,
                const int32_t *task_i, int n_task_i, const float *task_f, int n_task_f, B200Handle **out);
void b200_destroy(B200Handle *h);
int b200_set_terrain(B200Handle *h, const int16_t *dev_height_samples, int rows, int cols,
                     const float *dev_terrain_origins, int num_levels, int num_types);
int b200_bind_buffers(B200Handle *h, const B200Buffers *bufs);

/* clip + shift action history, save last_*, then `decimation` x (PD torque, rigid-body substep). */
int b200_dynamics_step(B200Handle *h, const float *dev_actions, void *cuda_stream);

/* Plugin mode (the stock LeggedRobot drives the backend through the Simulator API): GenesisSimulator.step only
 * (genesis_simulator.py:20-33) -- save last_*, then `decimation` x (PD torque, rigid-body substep) with `dev_actions`
 * taken as given.  LeggedRobot._pre_sim_step (legged_robot.py:230-252) has already clipped them, shifted its own action
 * history and applied its control-delay queue, so none of that is repeated here (B200Buffers.actions / last_actions /
 * llast_actions / action_queue are left untouched). */
int b200_simulator_step(B200Handle *h, const float *dev_actions, void *cuda_stream);

/* enabled = 0: warp slot w of the dynamics kernel runs env w.  Default (enabled, B200_DYN_ORDER=0 in the environment
 * disables it at creation): every b200_dynamics_step / b200_simulator_step also launches dynamics_order_kernel on an
 * internal side stream (forked before the dynamics launch, joined by the next b200_env_post_step), which sorts the envs
 * by the contact-solver work they needed two launches ago into B200Buffers.dyn_order so that the envs of a CTA cost the
 * same and the SMs get the same mix.  Pure scheduling: results are identical either way. */
int b200_set_dynamics_order(B200Handle *h, int enabled);
/* enabled = 0: dynamics_order_kernel is launched on the caller's stream instead of the internal side stream (profiling, or
 * callers that must keep every launch on one stream).  Default: enabled. */
int b200_set_side_stream(B200Handle *h, int enabled);

/* Fused post_physics_step. `step_counter` is LeggedRobot.common_step_counter *after* its increment;
 * `cmd_vx_lo/span` is the (curriculum-mutable) lin_vel_x command range as (lower, fp32(upper-lower)); `hist_step` is the number of
 * observation frames appended to the frame stacks so far (the caller counts the calls that ran PHASE_OBSERVE): this call's
 * frame goes to ring slots hist_step mod (K + 1) and hist_step mod (K + 1) + K + 1 of B200Buffers.obs_history / critic_obs. */
int b200_env_post_step(B200Handle *h, long long step_counter, float cmd_vx_lo, float cmd_vx_span, long long hist_step,
                       int phase_mask, void *cuda_stream);

/* One whole env.step (LeggedRobot.step, legged_robot.py:37-53) with HOST buffers, one call: takes `actions` ([N,A] fp32;
 * pinned host memory when `actions_on_host`, device memory otherwise) -- pinned memory is read by the kernels themselves
 * (mapped host memory: a copy kernel overlapped with the dynamics kernel's first phase, no copy-engine operation; the buffer
 * must stay untouched until the stream has passed the call, as with an asynchronous copy) --, then b200_dynamics_step and
 * b200_env_post_step(PHASE_ALL), then hands back rew_buf ([N] fp32),
 * reset_buf and time_out_buf ([N] bool bytes) into the given pinned host buffers (each may be NULL; when the three are one
 * 16-byte aligned slab rew | reset | time_out of at most 64 KB, laid out like the device side, the env kernel's last CTA writes
 * it into the mapped host memory itself, otherwise cudaMemcpyAsync does).  Environment switches, read at b200_create:
 * B200_ZERO_COPY_ACTIONS = 2 (default: staging copy kernel + programmatic dependent launch) | 1 (every env's warp reads its
 * actions in place) | 0 (copy engine); B200_ZERO_COPY_RESULTS = 1 (default) | 0 (copy engine).  Everything is
 * enqueued on `cuda_stream`; nothing is synchronised -- the caller waits on the stream before reading the host buffers.
 * This is the call the rollout loop makes (on_policy_runner.py:118-139: env.step(actions) followed by host reads). */
int b200_env_step(B200Handle *h, const float *actions, int actions_on_host, long long step_counter, float cmd_vx_lo, float cmd_vx_span,
                  long long hist_step, float *host_rew, uint8_t *host_reset, uint8_t *host_time_out, void *cuda_stream);

/* Device-stepped env.step: the same step as b200_env_step with device-resident actions, but every per-step scalar the
 * host passes there (step counter, frame-stack position, command range, go2_wtw behaviour ranges, the host-drawn sit-pose
 * coin and gait indices) lives in `dev_step_state`, B200_STEP_STATE_WORDS int32 words of caller-owned DEVICE memory, and is
 * moved forward by a one-thread kernel launched ahead of the dynamics kernel.  No launch parameter changes from one call to
 * the next, so a stream capture of T such calls -- with the policy network's kernels in between -- is a CUDA graph that
 * replays a whole rollout with one launch (on_policy_runner.py:118-139 is launch-bound in PyTorch eager mode).
 * Layout (B200_SS_*): [0] LeggedRobot.common_step_counter (the call increments it first, like legged_robot.py:57), [1]
 * observation frames appended so far (incremented too: the call runs with hist_step = the value it found), [2] cmd_vx_lo, [3] cmd_vx_span
 * (fp32 bit patterns), [4..11] the eight behaviour floats as b200_set_behavior stores them (lower, span pairs), [12]
 * sit-pose flag, [13] / [14] gait index of the callback / reset resampling (all three REdrawn by the call from the same
 * Philox stream the host uses, hcr_genesis_lr_cl_b200/host_rng.py), [15] number of gaits to draw from.  The caller writes
 * words 0..11 and 15 (and re-writes them when a curriculum moves a range); results are bit-identical to the host-stepped
 * calls with the same values.  `sit_init_percent` is the reference's cfg.init_state.sit_init_percent (0: no sit-pose resets).
 * The caller counts the steps itself (common_step_counter, hist_step) to keep its views of the frame-stack rings and of
 * the statistics ring in step; the bug-compatible R18 mode is not available here. */
#define B200_STEP_STATE_WORDS 16
#define B200_SS_STEP 0
#define B200_SS_HIST_STEP 1
#define B200_SS_VX_LO 2
#define B200_SS_VX_SPAN 3
#define B200_SS_BEH 4
#define B200_SS_SIT_POSE 12
#define B200_SS_GAIT_CB 13
#define B200_SS_GAIT_RESET 14
#define B200_SS_NUM_GAITS 15
int b200_env_step_device(B200Handle *h, const float *dev_actions, int32_t *dev_step_state, double sit_init_percent, void *cuda_stream);

/* Rollout-side fusion: where the NEXT fused post step (b200_env_post_step with PHASE_ALL / b200_env_step)
 * additionally writes its results, so that the runner's per-step copies disappear:
 *   rsl_rl/storage/rollout_storage.py:89-103   add_transitions: observations[t+1] / privileged_observations[t+1] <- obs, rewards[t],
 *                                              dones[t] -- the env kernel stores straight into those slabs;
 *   rsl_rl/algorithms/ppo.py:103-116           process_env_step: rewards += gamma * values * time_outs (bootstrapping on time-outs),
 *                                              applied in the kernel when `values` is given;
 *   rsl_rl/runners/on_policy_runner.py:127-136 per-step .cpu().numpy().tolist() episode book-keeping -> device counters:
 *                                              ep_return[env] += rew, ep_length[env] += 1; an env that resets adds both to
 *                                              ep_stats[0], ep_stats[1] and 1 to ep_stats[2], then restarts at zero.
 * All pointers are device pointers (any may be NULL); the targets apply to ONE step and are cleared by it. */
typedef struct B200RolloutTargets {
    float *obs;            /* [N, num_obs]   e.g. storage.observations[t + 1]           */
    float *privileged_obs; /* [N, num_priv]  e.g. storage.privileged_observations[t + 1] (estimator labels for the EE / DreamWaQ tasks) */
    float *next_state;     /* [N, num_obs]   go2_dreamwaq decoder target (next_state_buf), else NULL */
    float *rewards;        /* [N]            e.g. storage.rewards[t] ([N,1] contiguous)  */
    uint8_t *dones;        /* [N] uint8      e.g. storage.dones[t]                       */
    const float *values;   /* [N] critic values of the acting step (bootstrapping), or NULL */
    float gamma;
    float *ep_return;      /* [N] running return of the current episode   */
    float *ep_length;      /* [N] running length of the current episode   */
    float *ep_stats;       /* [3] sum of returns, sum of lengths, count of the episodes that ended (caller zeroes it when read) */
} B200RolloutTargets;
int b200_set_rollout_targets(B200Handle *h, const B200RolloutTargets *targets);

/* Per-step host scalars for the next b200_env_post_step / b200_reset_all: `sit_pose` != 0 -> envs that reset in that
 * call start in the sit pose (tron1_pf_ee.py:204-210 draws ONE coin per reset batch). */
int b200_set_step_flags(B200Handle *h, int sit_pose);

/* go2_wtw behaviour parameters (go2_wtw.py:180-217): the curriculum-mutable sampling ranges
 * {gait period, base height, foot clearance, pitch} x {lo, hi} and the gait index the host drew for the callback
 * resampling and for the reset resampling of the next step (ONE randint per call for all envs). */
int b200_set_behavior(B200Handle *h, const float *ranges8, int gait_callback, int gait_reset);

/* reset_idx(all envs) without a preceding step (BaseTask.reset). */
int b200_reset_all(B200Handle *h, long long step_counter, float cmd_vx_lo, float cmd_vx_span, void *cuda_stream);

/* static resource usage of a kernel ("dynamics" | "env"): registers/thread, static+dynamic smem/block, max blocks/SM */
int b200_kernel_info(B200Handle *h, const char *kernel, int *regs, int *smem_bytes, int *blocks_per_sm, int *block_threads);

/* which instantiation of the fused post_physics_step kernel this handle launches: the name of the built-in task preset
 * whose structural descriptor ints all equal the handle's (the kernel compiled with them as constants), or "generic"
 * (descriptor read at run time; any edited configuration).  Same results either way. */
const char *b200_env_kernel_variant(B200Handle *h);

/* number of kernel launches issued through this handle since creation */
long long b200_launch_count(B200Handle *h);

/* slots of the extras["episode"] ring inside B200Buffers.stats (a step's means live in slot step % ring) */
int b200_stats_ring(void);

/* CUDA device ordinal the handle lives on: the device that was current when b200_create ran.  Every entry point switches
 * to it for the duration of the call, so callers need not keep it current (legged_gym passes device strings only). */
int b200_device(B200Handle *h);

const char *b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_STEP_H */
