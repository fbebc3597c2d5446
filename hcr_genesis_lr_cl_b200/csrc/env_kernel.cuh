// env_kernel.cuh -- the fused post_physics_step: one launch, one pass over the per-env state in HBM.
//
// Replaces (reference paths relative to the LeggedGym-Ex root) for the `go2` and `go2_ts` observation layouts:
//   LeggedRobot.post_physics_step            legged_gym/envs/base/legged_robot.py:55-76
//   GenesisSimulator.post_physics_step       legged_gym/simulator/genesis_simulator.py:35-60
//     _check_base_pos_out_of_bound           :612-628      _update_surrounding_heights   :552-577
//     _calc_terrain_info_around_feet         :579-610
//   _post_physics_step_callback / _resample_commands / push_robots   legged_robot.py:300-334, genesis_simulator.py:150-158
//   check_termination                        legged_robot.py:78-92
//   compute_reward + _reward_*               legged_robot.py:150-168,458-608; go2_ts.py:133-176
//   reset_idx (+ curriculum, _reset_dofs, _reset_root_states, simulator.reset_idx DR, history clearing)
//                                            legged_robot.py:94-148,254-298; go2_ts.py:75-96; genesis_simulator.py:62-148,665-739;
//                                            legged_robot_ts.py:120-125
//   compute_observations (+ noise, history deques, clip)   go2.py:40-90; go2_ts.py:5-84; legged_robot_ts.py:59-76
//
// HBM-bound (~16.7 KB per env per policy step for go2_ts, dominated by the 900-float obs history and
// the 885-float critic stack).  One warp owns one env: every [N,k] row-major tensor row is read/written by
// consecutive lanes (coalesced), per-env scalars are computed redundantly by all lanes (no divergence), per-joint /
// per-foot / per-link quantities live one per lane and are reduced with shuffles.  Integer-valued outputs (cell
// indices, reset flags) follow the reference's fp32 operation order with explicit _rn intrinsics (no FMA
// contraction, true division by horizontal_scale; SURVEY "a-kernel notes").
#pragma once
#include "cuda_compat.cuh"
#include "philox.cuh"
#include "task_dev.cuh"          // TaskDev, TerrainDev
#include "env_presets.inc"       // generated: the int descriptor of every built-in task preset as compile-time tables

// The kernel is instantiated once per built-in preset with the structural ints of the task descriptor (widths, link
// lists, reward list, feature flags) as compile-time constants, and once generically.  `ti[k]` reads through this view:
// a constant for the specialised instantiations (except the run-time entries, env_ti_is_runtime), the descriptor else.
// 70 % of the generic kernel's instructions are integer/branch work on exactly these values.
struct EnvSpecGeneric {
    static constexpr bool fixed = false;
    __host__ __device__ static constexpr int get(int) { return 0; }
};
template <class S> struct TiView {
    const int *rt;
    __device__ __forceinline__ int operator[](int k) const { return (S::fixed && !env_ti_is_runtime(k)) ? S::get(k) : rt[k]; }
    __device__ __forceinline__ int dyn(int k) const { return rt[k]; }      // lane-indexed lists: always the run-time table
};

#ifndef ENV_WARPS_PER_BLOCK
#define ENV_WARPS_PER_BLOCK 4
#endif
#ifndef ENV_MIN_BLOCKS
#define ENV_MIN_BLOCKS 7
#endif
#define ES_OBS 0                   // clean obs [B200_MAX_OBS = 64]
#define ES_NOISY 64                // noisy obs [64]
#define ES_CRIT 128                 // single critic frame [<=192]
#define ES_PRIV 320                // privileged obs [<=112]
#define ES_MH 432                  // measured heights of this env [<=96]   (outputs that later phases re-read:
#define ES_LCS 528                 // link contact states [<=32]             kept on chip instead of a global round trip)
#define ES_HAF 560                 // height around feet [4*9]
#define ES_NV 596                  // terrain normals around feet [4*3]
#define ES_TOTAL 608
#define ENV_IN_WORDS 288           // staged per-env input rows (all small state tensors), words per env

#define ENV_STATS_RING 32           // per-step episode statistics are kept for this many policy steps

struct EnvCall {
    uint32_t step;       // LeggedRobot.common_step_counter after its increment
    float vx_lo, vx_span;
    uint32_t hist_step;  // observation frames appended to the frame stacks so far: this step's frame goes to ring slot hist_step mod K
    int phase_mask;
    int force_reset;     // b200_reset_all: run only the reset phase, for every env
    int sit_pose;        // envs resetting in this call start in the sit pose (one host coin per step, tron1_pf_ee.py:204-210)
    float beh[8];        // go2_wtw behaviour ranges {gait period, base height, foot clearance, pitch} x {lo, span}
    int gait_cb, gait_reset;   // gait index the host drew for the callback / reset resampling of this step
    B200RolloutTargets ro;   // rollout-side fusion: extra destinations of this step's results (all NULL: none)
    // extras["episode"] means: the last CTA to finish turns the per-step reductions into ring slot `stats_slot`
    int finalize, stats_slot;
    float inv_episode_length_s, inv_num_envs, inv_teacher, inv_student;
    // Device-stepped launches (b200_env_step_device): the per-step scalars above -- step, hist_step, vx_lo / vx_span, beh[],
    // sit_pose, gait_cb / gait_reset, stats_slot -- are NOT taken from this struct but from the device-resident step state
    // (include/b200_step.h, B200_SS_*) that step_advance_kernel moved forward just before, so that a captured CUDA graph of
    // the step can be replayed without any per-launch parameter from the host.  NULL for host-stepped launches.
    const int32_t *dstate;
    // b200_env_step with pinned host result buffers: device-visible address of the caller's packed host slab
    // (rew f32[N] | reset u8[N] | time_out u8[N], 16-byte aligned, 6 N a multiple of 16) -- the CTA that finalises the
    // step's statistics also copies the slab from rew_buf into it with 16-byte stores over PCIe, so that no copy-engine
    // operation follows the kernel.  NULL: nothing is written to the host by the kernel.
    uint4 *host_slab;
    int host_slab_vecs;
};

// The device-resident step state one policy step further (b200_env_step_device launches it ahead of the dynamics kernel):
// the counters, and the scalars the host would have drawn for this step from the Philox stream of env id 0xFFFFFFFF
// (host_rng.py: the sit-pose coin of tron1_pf_ee.py:204-210 and the two gait indices of go2_wtw.py:205-206), compared in
// integers exactly like the host compares its doubles.
#ifdef B200_ENV_DEFINE_KERNELS
__global__ void step_advance_kernel(int32_t *ss, uint32_t k0, uint32_t k1, uint32_t sit_threshold_q24, int behavior, int site_host) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t step = (uint32_t)ss[B200_SS_STEP] + 1u;
    ss[B200_SS_STEP] = (int32_t)step;
    ss[B200_SS_HIST_STEP] = ss[B200_SS_HIST_STEP] + 1;
    const Philox4 p = philox4x32_10(0xFFFFFFFFu, step, (uint32_t)site_host, 0u, k0, k1);
    ss[B200_SS_SIT_POSE] = (sit_threshold_q24 != 0u && (p.w[0] >> 8) < sit_threshold_q24) ? 1 : 0;
    if (behavior) {
        const uint32_t ng = (uint32_t)max(ss[B200_SS_NUM_GAITS], 1);
        ss[B200_SS_GAIT_CB] = (int32_t)min((uint32_t)(((uint64_t)(p.w[1] >> 8) * ng) >> 24), ng - 1u);
        ss[B200_SS_GAIT_RESET] = (int32_t)min((uint32_t)(((uint64_t)(p.w[2] >> 8) * ng) >> 24), ng - 1u);
    }
}
#else
__global__ void step_advance_kernel(int32_t *ss, uint32_t k0, uint32_t k1, uint32_t sit_threshold_q24, int behavior, int site_host);
#endif   // B200_ENV_DEFINE_KERNELS

// shared memory per CTA: header (mbarrier, last-CTA flag, the CTA's every-env statistics) + per warp (scratch, staged input rows)
#define ENV_HDR_WORDS 8             // [0,1] mbarrier  [2] last-CTA flag  [4] sum of terrain levels  [5] of the teacher envs  [6] sum of cstr_prob
#define ENV_HDR_ACC 4
__host__ __device__ inline int env_smem_bytes(int warps) { return ENV_HDR_WORDS * 4 + warps * (ES_TOTAL + ENV_IN_WORDS) * 4; }

// quat_rotate_inverse (math_utils.py:63-76), q = xyzw
__device__ __forceinline__ f3 rot_inv(float qx, float qy, float qz, float qw, f3 v) {
    const f3 qv = mk3(qx, qy, qz);
    const float s = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, qw), qw), 1.0f);
    const f3 a = mk3(__fmul_rn(v.x, s), __fmul_rn(v.y, s), __fmul_rn(v.z, s));
    const float w2 = __fmul_rn(2.0f, qw);
    const f3 cr = mk3(__fsub_rn(__fmul_rn(qv.y, v.z), __fmul_rn(qv.z, v.y)), __fsub_rn(__fmul_rn(qv.z, v.x), __fmul_rn(qv.x, v.z)),
                      __fsub_rn(__fmul_rn(qv.x, v.y), __fmul_rn(qv.y, v.x)));
    const f3 b = mk3(__fmul_rn(cr.x, w2), __fmul_rn(cr.y, w2), __fmul_rn(cr.z, w2));
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(qv.x, v.x), __fmul_rn(qv.y, v.y)), __fmul_rn(qv.z, v.z));
    const float d2 = __fmul_rn(2.0f, d);
    const f3 c = mk3(__fmul_rn(qv.x, d2), __fmul_rn(qv.y, d2), __fmul_rn(qv.z, d2));
    return mk3(__fadd_rn(__fsub_rn(a.x, b.x), c.x), __fadd_rn(__fsub_rn(a.y, b.y), c.y), __fadd_rn(__fsub_rn(a.z, b.z), c.z));
}

__device__ __forceinline__ float norm3_rn(float x, float y, float z) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// trunc((p + border) / hscale) as torch evaluates it in fp32 (genesis_simulator.py:564-565,582-583)
__device__ __forceinline__ int cell_of(float p, float border, float hscale) {
    return (int)truncf(__fdiv_rn(__fadd_rn(p, border), hscale));
}

__device__ __forceinline__ int wrap_idx(int i, int n) { return i < 0 ? i + n : i; }   // torch negative-index wrap

// scipy.stats.vonmises.cdf(x, kappa, loc) for kappa < 50, x - loc given: scipy's series (scipy/stats/_stats.pyx von_mises_cdf:
// shift x into one period, backward recursion over p - 1 terms, clip to [0, 1], add the period count back).  `tab` holds the
// kappa-only part of the recursion as pairs (R_n, R_n / n), n = p-1 .. 1 (TaskSpec.von_mises_series).  The reference evaluates
// this on the host in float64 (go2_wtw.py:423-429); fp32 here, ~1e-6 absolute.
__device__ __forceinline__ float von_mises_cdf(const float *tab, int p, float x) {
    const float ix = rintf(x * 0.15915494309189535f);
    x = fmaf(-ix, 6.2831853071795862f, x);
    float s, c, sn, cn;
    sincosf(x, &s, &c);
    sincosf((float)p * x, &sn, &cn);
    float V = 0.f;
    for (int n = 0; n < p - 1; n++) {
        const float sn2 = sn * c - cn * s;
        cn = cn * c + sn * s; sn = sn2;
        V = fmaf(tab[2 * n + 1], sn, tab[2 * n] * V);
    }
    const float F = fminf(fmaxf(0.5f + x * 0.15915494309189535f + V * 0.31830988618379069f, 0.f), 1.f);
    return F + ix;
}

// Frame stacks (obs_history: K = frame_stack frames of num_obs floats; critic stack: c_frame_stack frames) live in HBM as
// DOUBLE-WRITTEN RINGS of period M = K + 1: a row holds 2 M frame slots and the frame of observation step t is stored twice,
// in slot t mod M and in slot t mod M + M.  The K most recent frames, oldest first -- exactly the tensor the reference
// re-concatenates from its deque every step (legged_robot_ts.py:29-47) -- are then ALWAYS the contiguous slots
// [t mod M + 2, t mod M + M] of the row: the host hands them out as a strided view (row stride 2 M frames) and a step costs
// two frame writes per stack instead of moving the K - 1 kept frames (12.5 KB per env and step for go2_ts, 77 % of what the
// path used to touch).  The one spare slot per period is what keeps the view of step t intact while step t + 1 writes
// (slots t mod M + 1 and t mod M + 1 + M lie just outside it): rsl_rl's algorithms keep a reference to the tensors they
// acted on and copy them into the rollout storage only AFTER env.step (ppo.py:91-116), which the reference's freshly
// concatenated tensors allow.  A reset clears the env's whole row first (legged_robot_ts.py:120-125 zeroes every deque entry).
__device__ __forceinline__ void ring_clear(float *ring, int env, int K, int frame, int lane) {
    float *d = ring + (size_t)env * 2 * (K + 1) * frame;
    for (int e = lane; e < 2 * (K + 1) * frame; e += 32) d[e] = 0.f;
}
__device__ __forceinline__ void ring_put(float *ring, int env, int K, int frame, uint32_t hist_step, const float *newf, int lane) {
    const int M = K + 1, slot = (int)(hist_step % (uint32_t)M);
    float *d = ring + ((size_t)env * 2 * M + slot) * frame;
    for (int e = lane; e < frame; e += 32) { const float v = newf[e]; d[e] = v; d[(size_t)M * frame + e] = v; }
}

// per-env input tensors staged per CTA: X(field, element type, elements per env)
#define ENV_STAGED_INPUTS(X, A_, F_, L_, NS_)                                                                          \
    X(base_pos, float, 3) X(base_quat_wxyz, float, 4) X(base_lin_w, float, 3) X(base_ang_w, float, 3) X(env_origins, float, 3) \
    X(commands, float, 4) X(episode_length, int32_t, 1) X(fail_buf, int32_t, 1) X(terrain_levels, int64_t, 1)          \
    X(terrain_types, int64_t, 1) X(rand_push_vels, float, 3) X(dof_pos, float, A_) X(dof_vel, float, A_)               \
    X(actions, float, A_) X(last_actions, float, A_) X(llast_actions, float, A_) X(torques, float, A_)                 \
    X(last_dof_vel, float, A_) X(feet_pos, float, 3 * F_) X(feet_vel, float, 3 * F_) X(last_feet_vel, float, 3 * F_)   \
    X(link_contact_forces, float, 3 * L_) X(episode_sums, float, NS_) X(feet_air_time, float, F_)                      \
    X(last_contacts, uint8_t, F_) X(friction, float, 1) X(added_mass, float, 1) X(com_bias, float, 3)                  \
    X(kp_scale, float, A_) X(kd_scale, float, A_) X(gait_state, float, B200_GAIT_STATE)
#define ENV_MAX_STAGED 40
#define ENV_NOT_STAGED 0xFFFFFFFFu

// The per-env inputs as the warp code reads them, `R.field[row * k + i]`: in a staged CTA the pointers lead into the
// CTA's shared-memory slab that TMA filled and row = warp; otherwise they are the global tensors and row = env.
#define X_DECL(field, type, k) const type *field;
struct EnvInputs { ENV_STAGED_INPUTS(X_DECL, 0, 0, 0, 0) };
#undef X_DECL

// Byte offset of staged input `idx` inside a CTA's slab (ENV_NOT_STAGED when its CTA slice breaks the 16-byte rule of
// cp.async.bulk or the slab is full -- that tensor is then read straight from global memory), and the bytes one CTA
// brings in (idx = -1).  One definition for the host-built copy table and for the kernel: the specialised
// instantiations fold the offsets into the load instructions.
__host__ __device__ constexpr uint32_t env_stage_layout(int idx, int A, int F, int L, int NSUM, int nwarps) {
    uint32_t off = 0, total = 0, res = ENV_NOT_STAGED;
    int t = 0;
#define X_LAY(field, type, k) { const uint32_t b_ = (uint32_t)nwarps * (uint32_t)((k) * sizeof(type));                               \
        const bool st_ = b_ > 0 && (b_ & 15u) == 0u && off + b_ <= (uint32_t)(nwarps * ENV_IN_WORDS * 4);                            \
        if (t == idx) res = st_ ? off : ENV_NOT_STAGED;                                                                              \
        if (st_) { total += b_; off += (b_ + 15u) & ~15u; }                                                                          \
        t++; }
    ENV_STAGED_INPUTS(X_LAY, A, F, L, NSUM)
#undef X_LAY
    return idx < 0 ? total : res;
}

// Copy table, built once on the host when the buffers are bound: for every per-env input tensor its base pointer, bytes
// per env and byte offset inside a CTA's input slab.  Thread t of a CTA issues the bulk copy of entry t, so the
// per-thread set-up is a handful of instructions instead of the address arithmetic of all tensors.
struct EnvStageTab {
    const char *src[ENV_MAX_STAGED];
    uint32_t row_bytes[ENV_MAX_STAGED];
    uint32_t off[ENV_MAX_STAGED];
    int n;                 // entries
    uint32_t in_bytes;     // bytes the staged entries of one CTA bring in
    int ok;                // staging enabled
};

inline EnvStageTab env_stage_table(const TaskDev &T, const B200Buffers &B, int nwarps) {
    EnvStageTab tab;
    const int A = T.i[TI_A], F = T.i[TI_F], L = T.i[TI_L], NSUM = T.i[TI_N_SUMS];
    tab.n = 0; tab.ok = 1;
#define X_TAB(field, type, k) { tab.src[tab.n] = (const char *)B.field; tab.row_bytes[tab.n] = (uint32_t)((k) * sizeof(type));   \
        tab.off[tab.n] = env_stage_layout(tab.n, A, F, L, NSUM, nwarps); if (!B.field) tab.ok = 0; tab.n++; }
    ENV_STAGED_INPUTS(X_TAB, A, F, L, NSUM)
#undef X_TAB
    tab.in_bytes = env_stage_layout(-1, A, F, L, NSUM, nwarps);
    for (int k = tab.n; k < ENV_MAX_STAGED; k++) { tab.src[k] = nullptr; tab.row_bytes[k] = 0; tab.off[k] = ENV_NOT_STAGED; }
    return tab;
}

// `staged`: the CTA's inputs were brought into shared memory by bulk copies completing on the mbarrier `bar`.
// `R` / `row` is the view all per-env *inputs* are read through (EnvInputs): the global tensors with row = env, or (staged
// CTAs) the CTA's shared-memory slab that TMA filled with row = warp -- one exposed DRAM latency per CTA instead of one
// per dependent load.  All stores go to B (global memory).
template <class S>
// Must inline: it takes the kernel parameters by reference (an out-of-line call would copy them to the stack).
// Returns (thread 0 of a CTA-synchronous CTA only) the finalize ticket the CTA took once its statistics were out, else -1.
__device__ __forceinline__ int env_post_step_warp(const TaskDev &T, const B200Buffers &B, const EnvInputs &R, const int row, const TerrainDev &tr,
                                   const EnvCall &call, float *es, int env, int lane, bool staged, bool cta_sync, uint64_t *bar, float *cta_acc) {
    const float *tf = T.f;
    const TiView<S> ti{T.i};
    const int A = ti[TI_A], F = ti[TI_F], L = ti[TI_L], P = ti[TI_PX] * ti[TI_PY];
    const int n_sums = ti[TI_N_SUMS];
    const int pm = call.force_reset ? PHASE_RESET : call.phase_mask;
    EnvRng rng; rng.k0 = (uint32_t)ti[TI_SEED_LO]; rng.k1 = (uint32_t)ti[TI_SEED_HI]; rng.env = (uint32_t)(env + ti[TI_ENV_OFFSET]); rng.step = call.step;
    const float dt = tf[TF_POLICY_DT];
    if (staged) mbar_wait(bar, 0);                              // staged CTA: every input slab has landed
    // Staged CTAs hold only live envs, so their warps can meet at CTA barriers between the sections below: the warps of
    // an SM then walk the (several hundred KB of) straight-line code together and share instruction-cache lines instead
    // of each missing on its own (ncu: `no_instruction` was the second-largest stall).  Uniform per CTA by construction.
#ifndef ENV_NO_SECTION_SYNC
#define ENV_SECTION_SYNC() do { if (cta_sync) __syncthreads(); } while (0)
#else
#define ENV_SECTION_SYNC()
#endif

    // ------------------------------------------------------------------ per-env scalars (replicated in all lanes)
    f3 bp = mk3(R.base_pos[row * 3], R.base_pos[row * 3 + 1], R.base_pos[row * 3 + 2]);
    const float Qw = R.base_quat_wxyz[row * 4], Qx = R.base_quat_wxyz[row * 4 + 1], Qy = R.base_quat_wxyz[row * 4 + 2], Qz = R.base_quat_wxyz[row * 4 + 3];
    f3 vw = mk3(R.base_lin_w[row * 3], R.base_lin_w[row * 3 + 1], R.base_lin_w[row * 3 + 2]);
    f3 ww = mk3(R.base_ang_w[row * 3], R.base_ang_w[row * 3 + 1], R.base_ang_w[row * 3 + 2]);
    f3 origin = mk3(R.env_origins[row * 3], R.env_origins[row * 3 + 1], R.env_origins[row * 3 + 2]);
    float cmd0 = R.commands[row * 4], cmd1 = R.commands[row * 4 + 1], cmd2 = R.commands[row * 4 + 2], cmd3 = R.commands[row * 4 + 3];
    int ep_len = R.episode_length[row];
    int fail_cnt = R.fail_buf[row];
    // read before any lane may overwrite them below (lanes are not in lock-step between collectives)
    const int level0 = ti[TI_TERRAIN_CURRICULUM] ? (int)R.terrain_levels[row] : 0;
    const int ttype = ti[TI_TERRAIN_CURRICULUM] ? (int)R.terrain_types[row] : 0;
    f3 push_vel = mk3(R.rand_push_vels[row * 3], R.rand_push_vels[row * 3 + 1], 0.f);
    f3 lin_b, ang_b, grav;           // body-frame velocities, projected gravity
    float e_roll = 0.f, e_pitch = 0.f;   // base_euler (get_euler_xyz) for tracking_orientation
    if (ti[TI_BEHAVIOR]) {
        e_roll = atan2f(2.0f * (Qw * Qx + Qy * Qz), Qw * Qw - Qx * Qx - Qy * Qy + Qz * Qz);
        const float sinp = 2.0f * (Qw * Qy - Qz * Qx);
        e_pitch = fabsf(sinp) >= 1.f ? copysignf(1.5707963267948966f, sinp) : asinf(sinp);
    }
    // periodic-gait state (tron1_pf_ee.py:186-197): theta_left/right, gait time, phase
    float th0 = 0.f, th1 = 0.f, th2 = 0.f, th3 = 0.f, gtime = 0.f, gphi = 0.f, gper = tf[TF_GAIT_PERIOD], bh_t = 0.f, fc_t = 0.f, pt_t = 0.f, expc_frc = 0.f;
    const bool r18_env0 = ti[TI_R18] && (env + ti[TI_ENV_OFFSET]) == 0;
    const int r18_flags = r18_env0 ? (B.global_flags[0] >> 8) : 0;
    if (ti[TI_GAIT]) {
        const float *gs = R.gait_state + row * B200_GAIT_STATE;
        th0 = gs[B200_GS_TH]; th1 = gs[B200_GS_TH + 1]; th2 = gs[B200_GS_TH + 2]; th3 = gs[B200_GS_TH + 3]; gtime = gs[B200_GS_GT]; gphi = gs[B200_GS_PHI];
        if (ti[TI_BEHAVIOR]) { gper = gs[B200_GS_PER]; bh_t = gs[B200_GS_BH]; fc_t = gs[B200_GS_FC]; pt_t = gs[B200_GS_PT]; }
    }
    // go2_wtw.py:180-217: behaviour parameters of one env (the pronk/bound clearance clamp is applied per resampled env)
    auto resample_behavior = [&](int site, int gait) {
        gper = rand_range(call.beh[0], call.beh[1], rng.u(site, 0));
        bh_t = rand_range(call.beh[2], call.beh[3], rng.u(site, 1));
        fc_t = rand_range(call.beh[4], call.beh[5], rng.u(site, 2));
        pt_t = rand_range(call.beh[6], call.beh[7], rng.u(site, 3));
        const float *tl = tf + TF_GAIT_THETA + gait * B200_MAX_FEET;
        th0 = tl[0]; th1 = tl[1]; th2 = tl[2]; th3 = tl[3];
        if (th0 == 0.f && th1 == 0.f && ((th2 == 0.f && th3 == 0.f) || (th2 == 0.5f && th3 == 0.5f))) fc_t = call.beh[4];
    };
    // ------------------------------------------------------------------ per-lane values
    const bool jl = lane < A, fl = lane < F;
    float qj = jl ? R.dof_pos[row * A + lane] : 0.f, qdj = jl ? R.dof_vel[row * A + lane] : 0.f;
    float actj = jl ? R.actions[row * A + lane] : 0.f;
    const float q0j = jl ? tf[TF_DEFAULT_DOF_POS + lane] : 0.f;
    f3 fpos = mk3(0.f, 0.f, 0.f), fvel = mk3(0.f, 0.f, 0.f);
    if (fl) {
        const float *a = R.feet_pos + (row * F + lane) * 3, *b = R.feet_vel + (row * F + lane) * 3;
        fpos = mk3(a[0], a[1], a[2]); fvel = mk3(b[0], b[1], b[2]);
    }
    float ffz = fl ? R.link_contact_forces[(row * L + ti.dyn(TI_FEET_LINKS + lane)) * 3 + 2] : 0.f;   // foot contact force z
    float hmean = 0.f, hmax = 0.f;   // mean / max of the 9 terrain heights around this lane's foot

    if (pm & PHASE_CALLBACK) ep_len += 1;                       // legged_robot.py:60
    __syncwarp();            // every lane holds its copy of the state read above before any lane overwrites it (OOB teleport, resets)

    // ================================================================== GenesisSimulator.post_physics_step
    if (pm & PHASE_SIM_POST) {
        const bool oob = (bp.x >= tf[TF_X_HI]) || (bp.x <= tf[TF_X_LO]) || (bp.y >= tf[TF_Y_HI]) || (bp.y <= tf[TF_Y_LO]);
        if (oob) {                                               // genesis_simulator.py:612-628
            const f3 nb = mk3(__fadd_rn(tf[TF_INIT_POS], origin.x), __fadd_rn(tf[TF_INIT_POS + 1], origin.y), __fadd_rn(tf[TF_INIT_POS + 2], origin.z));
            fpos = fpos + (nb - bp);
            if (fl) {
                float *a = B.feet_pos + (env * F + lane) * 3, *b = const_cast<float *>(R.feet_pos) + (row * F + lane) * 3;
                a[0] = fpos.x; a[1] = fpos.y; a[2] = fpos.z; b[0] = fpos.x; b[1] = fpos.y; b[2] = fpos.z;
            }
            bp = nb;
            if (lane < 3) B.base_pos[env * 3 + lane] = comp3(bp, lane);
        }
        lin_b = rot_inv(Qx, Qy, Qz, Qw, vw);
        ang_b = rot_inv(Qx, Qy, Qz, Qw, ww);
        grav = rot_inv(Qx, Qy, Qz, Qw, mk3(0.f, 0.f, -1.f));
        if (lane < 3) {
            B.base_lin_vel[env * 3 + lane] = comp3(lin_b, lane);
            B.base_ang_vel[env * 3 + lane] = comp3(ang_b, lane);
            B.projected_gravity[env * 3 + lane] = comp3(grav, lane);
        }
        if (lane < 4) B.base_quat[env * 4 + lane] = lane == 0 ? Qx : (lane == 1 ? Qy : (lane == 2 ? Qz : Qw));
        if (lane == 0) {                                         // get_euler_xyz, math_utils.py:89-109
            const float sinr = 2.0f * (Qw * Qx + Qy * Qz), cosr = Qw * Qw - Qx * Qx - Qy * Qy + Qz * Qz;
            const float sinp = 2.0f * (Qw * Qy - Qz * Qx);
            const float siny = 2.0f * (Qw * Qz + Qx * Qy), cosy = Qw * Qw + Qx * Qx - Qy * Qy - Qz * Qz;
            B.base_euler[env * 3 + 0] = atan2f(sinr, cosr);
            B.base_euler[env * 3 + 1] = fabsf(sinp) >= 1.f ? copysignf(1.5707963267948966f, sinp) : asinf(sinp);
            B.base_euler[env * 3 + 2] = atan2f(siny, cosy);
        }
        if (ti[TI_CONTACT_STATES] && lane < ti[TI_N_CS]) {
            const float *f = R.link_contact_forces + (row * L + ti.dyn(TI_CS_LINKS + lane)) * 3;
            const float cs = norm3_rn(f[0], f[1], f[2]) > 1.0f ? 1.f : 0.f;
            B.link_contact_states[env * ti[TI_N_CS] + lane] = cs; es[ES_LCS + lane] = cs;
        }
        if (ti[TI_MEASURE_HEIGHTS]) {
            // quat_apply_yaw (math_utils.py:42-47) of the scan grid, then min of 3 int16 samples
            const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fmul_rn(Qz, Qz), __fmul_rn(Qw, Qw))), 1e-9f);
            const float zn = __fdiv_rn(Qz, nrm), wn = __fdiv_rn(Qw, nrm);
            for (int pt = lane; pt < P; pt += 32) {
                const float px = tf[TF_POINTS_X + pt / ti[TI_PY]], py = tf[TF_POINTS_Y + pt % ti[TI_PY]];
                const float tx = __fmul_rn(-__fmul_rn(zn, py), 2.0f), ty = __fmul_rn(__fmul_rn(zn, px), 2.0f);
                const float rx = __fadd_rn(__fadd_rn(px, __fmul_rn(wn, tx)), -__fmul_rn(zn, ty));
                const float ry = __fadd_rn(__fadd_rn(py, __fmul_rn(wn, ty)), __fmul_rn(zn, tx));
                int cx = cell_of(__fadd_rn(rx, bp.x), tf[TF_BORDER], tf[TF_HSCALE]);
                int cy = cell_of(__fadd_rn(ry, bp.y), tf[TF_BORDER], tf[TF_HSCALE]);
                cx = max(0, min(cx, tr.rows - 2)); cy = max(0, min(cy, tr.cols - 2));
                const int16_t *hp = tr.hf + (size_t)cx * tr.cols + cy;
                const int16_t h = min(min(__ldg(hp), __ldg(hp + tr.cols)), __ldg(hp + 1));
                const float mh = __fmul_rn((float)h, tf[TF_VSCALE]);
                B.measured_heights[env * P + pt] = mh; es[ES_MH + pt] = mh;
                if (B.height_cells) { B.height_cells[(env * P + pt) * 2] = cx; B.height_cells[(env * P + pt) * 2 + 1] = cy; }
            }
            if (ti[TI_FEET_INFO]) {
                int cx = cell_of(fpos.x, tf[TF_BORDER], tf[TF_HSCALE]), cy = cell_of(fpos.y, tf[TF_BORDER], tf[TF_HSCALE]);
                cx = max(0, min(cx, tr.rows - 2)); cy = max(0, min(cy, tr.cols - 2));
                if (fl) {
                    const int DX[9] = {-1, 1, 0, 0, 0, -1, 1, -1, 1}, DY[9] = {0, 0, -1, 1, 0, -1, 1, 1, -1};
                    int16_t hv[9];
                    float s = 0.f;
                    hmax = -3.0e38f;
#pragma unroll
                    for (int k = 0; k < 9; k++) {
                        hv[k] = __ldg(tr.hf + (size_t)wrap_idx(cx + DX[k], tr.rows) * tr.cols + wrap_idx(cy + DY[k], tr.cols));
                        const float hk = __fmul_rn((float)hv[k], tf[TF_VSCALE]);
                        B.height_around_feet[(env * F + lane) * 9 + k] = hk; es[ES_HAF + lane * 9 + k] = hk;
                        s = __fadd_rn(s, hk); hmax = fmaxf(hmax, hk);
                    }
                    hmean = __fdiv_rn(s, 9.0f);
                    const float den = __fmul_rn(tf[TF_HSCALE], 2.0f);
                    const float dx = __fdiv_rn((float)(int16_t)(hv[1] - hv[0]), den), dy = __fdiv_rn((float)(int16_t)(hv[3] - hv[2]), den);
                    const float nn = norm3_rn(dx, dy, -1.0f);
                    float *nv = B.normal_vector_around_feet + env * 3 * F + lane * 3;
                    nv[0] = __fdiv_rn(dx, nn); nv[1] = __fdiv_rn(dy, nn); nv[2] = __fdiv_rn(-1.0f, nn);
                    es[ES_NV + lane * 3] = nv[0]; es[ES_NV + lane * 3 + 1] = nv[1]; es[ES_NV + lane * 3 + 2] = nv[2];
                }
            }
        }
    } else {
        lin_b = mk3(B.base_lin_vel[env * 3], B.base_lin_vel[env * 3 + 1], B.base_lin_vel[env * 3 + 2]);
        ang_b = mk3(B.base_ang_vel[env * 3], B.base_ang_vel[env * 3 + 1], B.base_ang_vel[env * 3 + 2]);
        grav = mk3(B.projected_gravity[env * 3], B.projected_gravity[env * 3 + 1], B.projected_gravity[env * 3 + 2]);
        if (fl && ti[TI_FEET_INFO]) {
            float s = 0.f;
            hmax = -3.0e38f;
            for (int k = 0; k < 9; k++) { const float hk = B.height_around_feet[(env * F + lane) * 9 + k]; es[ES_HAF + lane * 9 + k] = hk; s = __fadd_rn(s, hk); hmax = fmaxf(hmax, hk); }
            hmean = __fdiv_rn(s, 9.0f);
            for (int k = 0; k < 3; k++) es[ES_NV + lane * 3 + k] = B.normal_vector_around_feet[env * 3 * F + lane * 3 + k];
        }
        if (ti[TI_MEASURE_HEIGHTS]) for (int pt = lane; pt < P; pt += 32) es[ES_MH + pt] = B.measured_heights[env * P + pt];
        if (ti[TI_CONTACT_STATES] && lane < ti[TI_N_CS]) es[ES_LCS + lane] = B.link_contact_states[env * ti[TI_N_CS] + lane];
    }
    __syncwarp();
    ENV_SECTION_SYNC();

    // ================================================================== _post_physics_step_callback
    if (pm & PHASE_CALLBACK) {
        if (ep_len % ti[TI_RESAMPLE_INTERVAL] == 0) {            // legged_robot.py:305-306, 317-334
            cmd0 = rand_range(call.vx_lo, call.vx_span, rng.u(SITE_CMD_RESAMPLE, 0));
            cmd1 = rand_range(tf[TF_CMD_VY_LO], tf[TF_CMD_VY_SPAN], rng.u(SITE_CMD_RESAMPLE, 1));
            if (ti[TI_HEADING_COMMAND]) cmd3 = rand_range(tf[TF_CMD_HEADING_LO], tf[TF_CMD_HEADING_SPAN], rng.u(SITE_CMD_RESAMPLE, 2));
            else cmd2 = rand_range(tf[TF_CMD_YAW_LO], tf[TF_CMD_YAW_SPAN], rng.u(SITE_CMD_RESAMPLE, 2));
            const float keep = norm3_rn(cmd0, cmd1, cmd2) > 0.2f ? 1.f : 0.f;
            cmd0 *= keep; cmd1 *= keep; cmd2 *= keep;
        }
        if (ti[TI_HEADING_COMMAND]) {                            // legged_robot.py:307-312
            // forward = quat_apply(base_quat, (1,0,0)); only x,y are needed
            const float t_y = __fmul_rn(Qz, 2.0f), t_z = __fmul_rn(-Qy, 2.0f);   // t = cross(q_xyz, e_x) * 2 = (0, 2 qz, -2 qy)
            const float fx = __fadd_rn(__fadd_rn(1.0f, __fmul_rn(Qw, 0.0f)), __fsub_rn(__fmul_rn(Qy, t_z), __fmul_rn(Qz, t_y)));
            const float fy = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(Qw, t_y)), __fsub_rn(__fmul_rn(Qz, 0.0f), __fmul_rn(Qx, t_z)));
            const float heading = atan2f(fy, fx);
            // wrap_to_pi as compiled by TorchScript: fmod (sign of the dividend), then -2pi where > pi  (quirk R16)
            float ang = fmodf(__fsub_rn(cmd3, heading), 6.2831853071795862f);
            ang = __fsub_rn(ang, ang > 3.1415926535897931f ? 6.2831853071795862f : 0.0f);
            cmd2 = fminf(fmaxf(__fmul_rn(0.5f, ang), tf[TF_CMD_YAW_LO]), tf[TF_CMD_YAW_HI]);
        }
        if (ti[TI_PUSH_ROBOTS] && (call.step % (uint32_t)ti[TI_PUSH_INTERVAL]) == 0) {   // genesis_simulator.py:150-158
            const float span = __fmul_rn(2.0f, tf[TF_MAX_PUSH]);
            push_vel.x = rand_range(-tf[TF_MAX_PUSH], span, rng.u(SITE_PUSH, 0));
            push_vel.y = rand_range(-tf[TF_MAX_PUSH], span, rng.u(SITE_PUSH, 1));
            vw.x = __fadd_rn(vw.x, push_vel.x); vw.y = __fadd_rn(vw.y, push_vel.y);
            if (lane < 2) { B.rand_push_vels[env * 3 + lane] = lane == 0 ? push_vel.x : push_vel.y; B.base_lin_w[env * 3 + lane] = lane == 0 ? vw.x : vw.y; }
        }
    }

    if ((pm & PHASE_CALLBACK) && ti[TI_BEHAVIOR] && ep_len % ti[TI_BEHAVIOR_INTERVAL] == 0) resample_behavior(SITE_BEHAVIOR, call.gait_cb);   // go2_wtw.py:258-263
    // ================================================================== check_termination
    bool time_out = false, reset = false;
    if (pm & PHASE_TERMINATION) {
        bool hit = false;
        if (lane < ti[TI_N_TERM]) {
            const float *f = R.link_contact_forces + (row * L + ti.dyn(TI_TERM_LINKS + lane)) * 3;
            hit = norm3_rn(f[0], f[1], f[2]) > 10.0f;
        }
        const bool fail = (__ballot_sync(B200_FULL_MASK, hit) != 0u) || (grav.z > tf[TF_MAX_PROJ_GRAV]);
        fail_cnt += fail ? 1 : 0;
        time_out = ep_len > ti[TI_MAX_EPISODE_LENGTH];
        reset = ((float)fail_cnt > tf[TF_FAIL_LIMIT]) || time_out;
        const bool refused = B.nonfinite != nullptr && B.nonfinite[env] != 0;
        __syncwarp();                                            // every lane has read the flag before lane 0 clears it
        if (refused) {                                           // the dynamics kernel refused a non-finite state: reset now (a failure, not a time-out)
            reset = true;
            if (lane == 0) { B.nonfinite[env] = 0; atomicAdd(B.global_flags + 2, 1); }
        }
    }
    if (call.force_reset) reset = true;

    // per-joint / per-foot history values used by the constraints and the rewards
    float lastj = 0.f, llastj = 0.f, tauj = 0.f, lqdj = 0.f;
    if ((pm & PHASE_REWARD) && jl) {
        lastj = R.last_actions[row * A + lane]; llastj = R.llast_actions[row * A + lane];
        tauj = R.torques[row * A + lane]; lqdj = R.last_dof_vel[row * A + lane];
    }
    // ================================================================== compute_reward
    float my_sum = (lane < n_sums) ? R.episode_sums[row * n_sums + lane] : 0.f;
    float fat = fl ? R.feet_air_time[row * F + lane] : 0.f;
    float cprob = 0.f;
    if ((pm & PHASE_REWARD) && ti[TI_CAT]) {
        // ============================================================== compute_constraints_cat (go2_cat.py:135-215)
        // violations are 0/1 and ConstraintManager's running maximum never exceeds 1 (constraint_manager.py:43-64), so
        // each constraint contributes max_p where violated; cstr_prob = max over the nine constraints.
        const bool c_torque = __ballot_sync(B200_FULL_MASK, jl && fabsf(tauj) > tf[TF_TORQUE_LIMIT + lane]) != 0u;
        const bool c_dofvel = __ballot_sync(B200_FULL_MASK, jl && fabsf(qdj) > tf[TF_DOF_VEL_LIMIT + lane]) != 0u;
        const bool c_arate = __ballot_sync(B200_FULL_MASK, jl && __fdiv_rn(fabsf(__fsub_rn(actj, lastj)), dt) > tf[TF_CAT_ACTION_RATE]) != 0u;
        float hs = 0.f;
        for (int pt = lane; pt < P; pt += 32) hs += bp.z - (ti[TI_MEASURE_HEIGHTS] ? es[ES_MH + pt] : 0.f);
        const bool c_height = warp_sum(hs) / (float)P < tf[TF_CAT_MIN_BASE_HEIGHT];
        bool hit = false;
        if (lane < ti[TI_N_PEN]) { const float *f = R.link_contact_forces + (row * L + ti.dyn(TI_PEN_LINKS + lane)) * 3; hit = norm3_rn(f[0], f[1], f[2]) > 10.0f; }
        const bool c_coll = __ballot_sync(B200_FULL_MASK, hit) != 0u;
        bool stumble = false;
        if (fl) { const float *f = R.link_contact_forces + (row * L + ti.dyn(TI_FEET_LINKS + lane)) * 3; stumble = norm3_rn(f[0], f[1], f[2]) > __fmul_rn(4.0f, fabsf(f[2])); }
        const bool c_stumble = __ballot_sync(B200_FULL_MASK, stumble) != 0u;
        const bool below = __ballot_sync(B200_FULL_MASK, jl && qj < tf[TF_DOF_LIM_LO + lane]) != 0u;
        const bool above = __ballot_sync(B200_FULL_MASK, jl && qj > tf[TF_DOF_LIM_HI + lane]) != 0u;
        const bool c_dofpos = below && above;                                 // product of two any() as written (go2_cat.py:168-169)
        const bool c_orient = grav.z > tf[TF_CAT_MAX_PROJ_GRAV];
        const bool fast_here = __ballot_sync(B200_FULL_MASK, jl && fabsf(qdj) > 4.0f) != 0u;
        const bool fast = ti[TI_CAT_GLOBAL_STANDSTILL] ? ((B.global_flags[0] & 1) != 0) : fast_here;
        const bool c_still = (norm3_rn(cmd0, cmd1, cmd2) < 0.1f) && fast;
        const float sp = tf[TF_CAT_SOFT_P];
        const bool cv[9] = {c_torque, c_dofvel, c_arate, c_height, c_coll, c_stumble, c_dofpos, c_orient, c_still};
        const float cp[9] = {sp, sp, sp, sp, 1.f, 1.f, 1.f, 1.f, sp};
        const int cbase = n_sums - 9;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            if (cv[k]) cprob = fmaxf(cprob, cp[k]);
            if (lane == cbase + k && cv[k]) my_sum = __fadd_rn(my_sum, 1.0f);
        }
        if (lane == 0) { B.cstr_prob[env] = cprob; if (cta_sync) atomicAdd(cta_acc + 2, cprob); else atomicAdd(B.stats + n_sums + 2, cprob); }
    }
    if (pm & PHASE_REWARD) {
        const float small_cmd = norm3_rn(cmd0, cmd1, cmd2) < 0.1f ? 1.f : 0.f;
        const float dqj = qj - q0j;
        float rew = 0.f;
        const int nr = ti[TI_N_REWARDS];
        // specialised instantiations unroll the loop: ids are constants, the switch folds to the active terms in order
        constexpr int kRewardUnroll = S::fixed ? B200_MAX_REWARDS : 1;
#pragma unroll kRewardUnroll
        for (int i = 0; i < nr; i++) {
            const int id = ti[TI_REWARD_IDS + i];
            float r = 0.f;
            ENV_SECTION_SYNC();
            switch (id) {
            case RW_ACTION_RATE: { const float d = lastj - actj; r = warp_sum(d * d); break; }
            case RW_ACTION_SMOOTHNESS: { const float d = actj - 2.0f * lastj + llastj; r = warp_sum(d * d); break; }
            case RW_ANG_VEL_XY: r = ang_b.x * ang_b.x + ang_b.y * ang_b.y; break;
            case RW_BASE_HEIGHT: {
                float s = 0.f;
                for (int pt = lane; pt < P; pt += 32) s += bp.z - (ti[TI_MEASURE_HEIGHTS] ? es[ES_MH + pt] : 0.f);
                const float m = warp_sum(s) / (float)P - tf[TF_BASE_HEIGHT_TARGET];
                r = m * m; break; }
            case RW_BIPED_PERIODIC_GAIT:                          // tron1_pf_ee.py:335-437 / go2_wtw.py:374-478, "step" indicator
            case RW_QUAD_PERIODIC_GAIT: {
                const int nf = id == RW_BIPED_PERIODIC_GAIT ? 2 : 4;
                float term = 0.f;
                if (lane < nf) {
                    const float *f = R.link_contact_forces + (row * L + ti.dyn(TI_FEET_LINKS + lane)) * 3;
                    const float q_frc = norm3_rn(f[0], f[1], f[2]), q_spd = norm3_rn(fvel.x, fvel.y, fvel.z);
                    const float thl = lane == 0 ? th0 : (lane == 1 ? th1 : (lane == 2 ? th2 : th3));
                    const float ph = __fmul_rn(fmodf(__fadd_rn(gphi, thl), 1.0f), 6.2831853071795862f);
                    const bool swing = ph >= 0.f && ph < tf[TF_GAIT_B_SWING], stance = ph >= tf[TF_GAIT_B_SWING] && ph < 6.2831853071795862f;
                    float expc_spd;
                    if (ti[TI_VM_TERMS] > 0) {                    // "smooth" indicator (go2_wtw.py:415-453, tron1_pf_ee.py:369-407)
                        const float *vm = tf + TF_VM_R;
                        const int pv = ti[TI_VM_TERMS];
                        const float FA = fminf(fmaxf(von_mises_cdf(vm, pv, ph), 0.f), 1.f);                                   // loc = a_swing = 0
                        const float FB = fminf(fmaxf(von_mises_cdf(vm, pv, ph - tf[TF_GAIT_B_SWING]), 0.f), 1.f);              // loc = b_swing
                        const float FS = fminf(fmaxf(von_mises_cdf(vm, pv, ph - 6.2831853071795862f), 0.f), 1.f);              // loc = b_stance = 2 pi
                        const float sw_ind = FA * (1.f - FB), st_ind = FB * (1.f - FS);
                        // outside the swing window: C_spd = -stance, C_frc = -1 + stance; inside: C_frc = -swing, C_spd = -1 + swing
                        const bool in_sw = swing || (r18_env0 && (r18_flags & (2 << lane)));      // R18: row 0 joins the swing set of any env
                        expc_spd = in_sw ? -1.f + sw_ind : -st_ind;
                        expc_frc = in_sw ? -sw_ind : -1.f + st_ind;
                        if (tf[TF_GAIT_B_SWING] == 0.f) { expc_frc = 0.f; expc_spd = -1.f; }                                  // standing gait
                    } else {
                        expc_frc = swing ? -1.f : 0.f;
                        expc_spd = stance ? -1.f : 0.f;
                        if (r18_env0) {                           // R18: row 0 rides along with "any env in swing", then "any env in stance"
                            if (r18_flags & (2 << lane)) { expc_frc = -1.f; expc_spd = 0.f; }
                            if (r18_flags & (32 << lane)) { expc_frc = 0.f; expc_spd = -1.f; }
                        }
                    }
                    term = __fadd_rn(__fmul_rn(expc_spd, q_spd), __fmul_rn(expc_frc, q_frc));
                }
                r = expf(warp_sum(term)); break; }
            case RW_COLLISION: {
                float hitf = 0.f;
                if (lane < ti[TI_N_PEN]) { const float *f = R.link_contact_forces + (row * L + ti.dyn(TI_PEN_LINKS + lane)) * 3; hitf = norm3_rn(f[0], f[1], f[2]) > 0.1f ? 1.f : 0.f; }
                r = warp_sum(hitf); break; }
            case RW_DOF_ACC: { const float d = (lqdj - qdj) / dt; r = warp_sum(d * d); break; }
            case RW_DOF_CLOSE_TO_DEFAULT: r = warp_sum(dqj * dqj); break;
            case RW_DOF_POS_LIMITS: {
                float o = 0.f;
                if (jl) o = -fminf(qj - tf[TF_DOF_LIM_LO + lane], 0.f) + fmaxf(qj - tf[TF_DOF_LIM_HI + lane], 0.f);
                r = warp_sum(o); break; }
            case RW_DOF_POS_STAND_STILL: r = warp_sum(dqj * dqj) * small_cmd; break;
            case RW_DOF_POWER: r = warp_sum(fabsf(tauj * qdj)); break;
            case RW_DOF_VEL: r = warp_sum(qdj * qdj); break;
            case RW_DOF_VEL_STAND_STILL: r = warp_sum(fabsf(qdj)) * small_cmd; break;
            case RW_FEET_AIR_TIME: {                              // stateful; go2_ts.py:133-145
                const bool contact = ffz > 1.0f;
                const bool lastc = fl ? (R.last_contacts[row * F + lane] != 0) : false;
                const bool filt = contact || lastc;
                if (fl) B.last_contacts[env * F + lane] = contact ? 1 : 0;
                const bool first = (fat > 0.f) && filt;
                fat = __fadd_rn(fat, dt);
                float s = warp_sum((fl && first) ? __fsub_rn(fat, tf[TF_AIR_TIME_THRESHOLD]) : 0.f);
                s *= norm3_rn(cmd0, cmd1, 0.f) > 0.1f ? 1.f : 0.f;
                if (filt) fat = 0.f;
                r = s; break; }
            case RW_FEET_CONTACT_STAND_STILL: {
                const float cnt = warp_sum((fl && ffz > 0.1f) ? 1.f : 0.f);
                r = (cnt == (float)F ? 1.f : 0.f) * small_cmd; break; }
            case RW_FEET_DISTANCE: {                              // tron1_pf.py:146-151 (feet 0 and 1)
                const float x0 = __shfl_sync(B200_FULL_MASK, fpos.x, 0), y0 = __shfl_sync(B200_FULL_MASK, fpos.y, 0);
                const float x1 = __shfl_sync(B200_FULL_MASK, fpos.x, 1), y1 = __shfl_sync(B200_FULL_MASK, fpos.y, 1);
                const float dx = x0 - x1, dy = y0 - y1;
                r = fmaxf(0.f, tf[TF_FOOT_DISTANCE_THRESHOLD] - sqrtf(dx * dx + dy * dy)); break; }
            case RW_NO_FLY: {                                     // tron1_pf.py:153-156
                const float cnt = warp_sum((fl && ffz > 0.1f) ? 1.f : 0.f);
                r = cnt == 1.f ? 1.f : 0.f; break; }
            case RW_FOOT_ACC: {
                float s = 0.f;
                if (fl) { const float *lv = R.last_feet_vel + (row * F + lane) * 3;
                    const float ax = (fvel.x - lv[0]) / dt, ay = (fvel.y - lv[1]) / dt, az = (fvel.z - lv[2]) / dt; s = ax * ax + ay * ay + az * az; }
                r = warp_sum(s); break; }
            case RW_FOOT_CLEARANCE: {                             // legged_robot.py:575-588; go2_ts.py:147-161
                float s = 0.f;
                if (fl) {
                    const float vxy = sqrtf(fvel.x * fvel.x + fvel.y * fvel.y);
                    float z = fpos.z;
                    if (ti[TI_CLEARANCE_MODE] == 1) z -= hmean; else if (ti[TI_CLEARANCE_MODE] == 2) z -= hmax;
                    const float e = z - tf[TF_FOOT_CLEARANCE_TARGET] - tf[TF_FOOT_HEIGHT_OFFSET];
                    s = vxy * e * e;
                }
                r = expf(-warp_sum(s) / tf[TF_FOOT_CLEARANCE_SIGMA]); break; }
            case RW_FOOT_LANDING_VEL: {
                float s = 0.f;
                if (fl) { const bool about = ((fpos.z - tf[TF_FOOT_HEIGHT_OFFSET]) < tf[TF_ABOUT_LANDING]) && !(ffz > 0.1f) && (fvel.z < 0.f); s = about ? fvel.z * fvel.z : 0.f; }
                r = warp_sum(s); break; }
            case RW_HIP_POS: r = warp_sum((jl && (lane % 3) == 0) ? dqj * dqj : 0.f); break;
            case RW_KEEP_BALANCE: r = 1.f; break;
            case RW_LIN_VEL_Z: r = lin_b.z * lin_b.z; break;
            case RW_ORIENTATION: r = grav.x * grav.x + grav.y * grav.y; break;
            case RW_THIGH_POS: r = warp_sum((jl && (lane % 3) == 1) ? dqj * dqj : 0.f); break;
            case RW_TORQUES: r = warp_sum(tauj * tauj); break;
            case RW_TRACKING_ANG_VEL: { const float e = cmd2 - ang_b.z; r = expf(-(e * e) / tf[TF_TRACKING_SIGMA]); break; }
            case RW_TRACKING_BASE_HEIGHT: {                       // tron1_pf_ee.py:439-444
                float s = 0.f;
                for (int pt = lane; pt < P; pt += 32) s += bp.z - (ti[TI_MEASURE_HEIGHTS] ? es[ES_MH + pt] : 0.f);
                const float m = warp_sum(s) / (float)P - (ti[TI_BEHAVIOR] ? bh_t : tf[TF_BASE_HEIGHT_TARGET]);
                r = expf(-(m * m) / tf[TF_BASE_HEIGHT_SIGMA]); break; }
            case RW_TRACKING_FOOT_CLEARANCE: {                    // go2_wtw.py:502-518
                float s = 0.f;
                if (fl) { const float e = fpos.z - fc_t - tf[TF_FOOT_HEIGHT_OFFSET]; s = sqrtf(fvel.x * fvel.x + fvel.y * fvel.y) * e * e; }
                r = expf(-warp_sum(s) / tf[TF_FOOT_CLEARANCE_SIGMA]); break; }
            case RW_TRACKING_ORIENTATION: { const float ep = e_pitch - pt_t; r = expf(-(e_roll * e_roll + ep * ep) / tf[TF_EULER_SIGMA]); break; }   // go2_wtw.py:497-500
            case RW_TRACKING_LIN_VEL: { const float ex = cmd0 - lin_b.x, ey = cmd1 - lin_b.y; r = expf(-(ex * ex + ey * ey) / tf[TF_TRACKING_SIGMA]); break; }
            default: break;
            }
            const float sr = __fmul_rn(r, tf[TF_REWARD_SCALE + id]);
            rew = __fadd_rn(rew, sr);
            if (lane == i) my_sum = __fadd_rn(my_sum, sr);
        }
        if (ti[TI_ONLY_POSITIVE]) {
            if (ti[TI_CAT]) rew = __fmul_rn(rew, __fsub_rn(1.0f, cprob));      // go2_cat.py:229-231
            rew = fmaxf(rew, 0.f);
        }
        if (ti[TI_TERMINATION_COL] >= 0) {
            const float sr = __fmul_rn((reset && !time_out) ? 1.f : 0.f, tf[TF_REWARD_SCALE + RW_TERMINATION]);
            rew = __fadd_rn(rew, sr);
            if (lane == ti[TI_TERMINATION_COL]) my_sum = __fadd_rn(my_sum, sr);
        }
        if (lane == 0) {
            B.rew_buf[env] = rew;
            if (call.ro.rewards)               // ppo.py:107-109: bootstrapping on time-outs, with the value of the acting step
                call.ro.rewards[env] = (call.ro.values && time_out) ? __fadd_rn(rew, __fmul_rn(call.ro.gamma, call.ro.values[env])) : rew;
            if (call.ro.ep_return) {           // on_policy_runner.py:127-136 on the device
                const float ret = call.ro.ep_return[env] + rew, len = call.ro.ep_length[env] + 1.f;
                if (reset) { atomicAdd(call.ro.ep_stats, ret); atomicAdd(call.ro.ep_stats + 1, len); atomicAdd(call.ro.ep_stats + 2, 1.f); }
                call.ro.ep_return[env] = reset ? 0.f : ret; call.ro.ep_length[env] = reset ? 0.f : len;
            }
        }
    }
    if ((pm & PHASE_TERMINATION) && lane == 0) {
        B.reset_buf[env] = reset ? 1 : 0; B.time_out_buf[env] = time_out ? 1 : 0;
        if (call.ro.dones) call.ro.dones[env] = reset ? 1 : 0;
    }

    if (ti[TI_GAIT] && (pm & PHASE_REWARD)) {                    // tron1_pf_ee.py:27-35: advance the gait clock after the reward
        gtime = __fadd_rn(gtime, dt);
        if (gtime >= __fsub_rn(gper, __fmul_rn(dt, 0.5f))) gtime = 0.f;
        if (r18_env0 && (r18_flags & 1)) gtime = 0.f;           // R18: row 0 is zeroed whenever any env wraps
        gphi = __fdiv_rn(gtime, gper);
    }
    ENV_SECTION_SYNC();
    // ================================================================== reset_idx
    f3 grav_obs = grav;
    int new_level = level0;
    if ((pm & PHASE_RESET) && reset) {
        if (ti[TI_OBS_KIND] >= 1) {                                               // history deques cleared (legged_robot_ts.py:120-125)
            ring_clear(B.obs_history, env, ti[TI_FRAME_STACK], ti[TI_NUM_OBS], lane);
            ring_clear(B.critic_obs, env, ti[TI_C_FRAME_STACK], ti[TI_SINGLE_CRITIC], lane);
            __syncwarp();                                                         // before the observation phase writes the new frame
        }
        if (lane < n_sums) atomicAdd(B.stats + lane, my_sum);                    // extras["episode"] numerators
        if (lane == 0) atomicAdd(B.stats + n_sums, 1.0f);
        my_sum = 0.f;
        if (ti[TI_TERRAIN_CURRICULUM]) {                                          // legged_robot.py:254-272; genesis_simulator.py:140-148
            const float ddx = __fsub_rn(bp.x, origin.x), ddy = __fsub_rn(bp.y, origin.y);
            const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));
            const bool up = dist > tf[TF_TERRAIN_HALF_LENGTH];
            const float cn = __fsqrt_rn(__fadd_rn(__fmul_rn(cmd0, cmd0), __fmul_rn(cmd1, cmd1)));
            const bool down = (dist < __fmul_rn(__fmul_rn(cn, tf[TF_EPISODE_LENGTH_S]), 0.5f)) && !up;
            int lv = level0 + (up ? 1 : 0) - (down ? 1 : 0);
            if (call.force_reset) lv = level0;               // init_done == False: no curriculum move
            const int nl = ti[TI_NUM_LEVELS];
            const int rnd = min((int)__fmul_rn(rng.u(SITE_LEVEL, 0), (float)nl), nl - 1);
            lv = lv >= nl ? rnd : max(lv, 0);
            const int ty = ttype;
            const float *og = tr.origins + ((size_t)lv * tr.types + ty) * 3;
            origin = mk3(og[0], og[1], og[2]);
            if (lane == 0) B.terrain_levels[env] = lv;
            new_level = lv;
            if (lane < 3) B.env_origins[env * 3 + lane] = comp3(origin, lane);
        }
        if (ti[TI_BEHAVIOR]) resample_behavior(SITE_BEHAVIOR_RESET, call.gait_reset);   // go2_wtw.py:124-126, before the commands
        {   // _resample_commands(env_ids)
            cmd0 = rand_range(call.vx_lo, call.vx_span, rng.u(SITE_CMD_RESET, 0));
            cmd1 = rand_range(tf[TF_CMD_VY_LO], tf[TF_CMD_VY_SPAN], rng.u(SITE_CMD_RESET, 1));
            if (ti[TI_HEADING_COMMAND]) cmd3 = rand_range(tf[TF_CMD_HEADING_LO], tf[TF_CMD_HEADING_SPAN], rng.u(SITE_CMD_RESET, 2));
            else cmd2 = rand_range(tf[TF_CMD_YAW_LO], tf[TF_CMD_YAW_SPAN], rng.u(SITE_CMD_RESET, 2));
            const float keep = norm3_rn(cmd0, cmd1, cmd2) > 0.2f ? 1.f : 0.f;
            cmd0 *= keep; cmd1 *= keep; cmd2 *= keep;
        }
        if (jl) {                                                                 // _reset_dofs: q0 + U(-r, r), qd = 0
            const float rr = tf[TF_RESET_DOF_NOISE + lane];
            qj = call.sit_pose ? tf[TF_SIT_DOF_POS + lane] : __fadd_rn(q0j, rand_range(-rr, __fmul_rn(2.0f, rr), rng.u(SITE_DOF, lane)));
            qdj = 0.f; actj = 0.f;
            B.dof_pos[env * A + lane] = qj; B.dof_vel[env * A + lane] = 0.f;
            B.actions[env * A + lane] = 0.f; B.last_actions[env * A + lane] = 0.f; B.llast_actions[env * A + lane] = 0.f;
            B.last_dof_vel[env * A + lane] = 0.f;
        }
        {   // _reset_root_states
            const int ip = call.sit_pose ? TF_SIT_POS : TF_INIT_POS;     // tron1_pf_ee.py:204-210,287-305
            bp = mk3(__fadd_rn(tf[ip], origin.x), __fadd_rn(tf[ip + 1], origin.y), __fadd_rn(tf[ip + 2], origin.z));
            if (ti[TI_HEIGHTFIELD]) {
                const float sp = __fmul_rn(2.0f, tf[TF_RESET_ROOT_XY]);
                bp.x = __fadd_rn(bp.x, rand_range(-tf[TF_RESET_ROOT_XY], sp, rng.u(SITE_ROOT, 0)));
                bp.y = __fadd_rn(bp.y, rand_range(-tf[TF_RESET_ROOT_XY], sp, rng.u(SITE_ROOT, 1)));
            }
            const float sv = __fmul_rn(2.0f, tf[TF_RESET_ROOT_VEL]), lo = -tf[TF_RESET_ROOT_VEL];
            lin_b = mk3(rand_range(lo, sv, rng.u(SITE_ROOT, 2)), rand_range(lo, sv, rng.u(SITE_ROOT, 3)), rand_range(lo, sv, rng.u(SITE_ROOT, 4)));
            ang_b = mk3(rand_range(lo, sv, rng.u(SITE_ROOT, 5)), rand_range(lo, sv, rng.u(SITE_ROOT, 6)), rand_range(lo, sv, rng.u(SITE_ROOT, 7)));
            if (call.sit_pose) { lin_b = mk3(0.f, 0.f, 0.f); ang_b = mk3(0.f, 0.f, 0.f); }
            const int iq = call.sit_pose ? TF_SIT_QUAT : TF_INIT_QUAT;
            const float ix = tf[iq], iy = tf[iq + 1], iz = tf[iq + 2], iw = tf[iq + 3];
            grav_obs = rot_inv(ix, iy, iz, iw, mk3(0.f, 0.f, -1.f));
            if (lane < 3) {
                B.base_pos[env * 3 + lane] = comp3(bp, lane);
                B.base_lin_w[env * 3 + lane] = comp3(lin_b, lane); B.base_ang_w[env * 3 + lane] = comp3(ang_b, lane);
                B.base_lin_vel[env * 3 + lane] = comp3(lin_b, lane); B.base_ang_vel[env * 3 + lane] = comp3(ang_b, lane);
                B.projected_gravity[env * 3 + lane] = comp3(grav_obs, lane);
                B.last_base_lin_vel[env * 3 + lane] = 0.f; B.last_base_ang_vel[env * 3 + lane] = 0.f;
            }
            if (lane < 4) {
                B.base_quat_wxyz[env * 4 + lane] = lane == 0 ? iw : (lane == 1 ? ix : (lane == 2 ? iy : iz));
                B.base_quat[env * 4 + lane] = lane == 0 ? ix : (lane == 1 ? iy : (lane == 2 ? iz : iw));
            }
        }
        // domain randomisation (genesis_simulator.py:62-82,665-739)
        // (stored to global memory and to the view the observation phase reads them back through)
#define DR_PUT(field, k, i, val) do { const float v_ = (val); B.field[env * (k) + (i)] = v_; const_cast<float *>(R.field)[row * (k) + (i)] = v_; } while (0)
        if (lane == 0) {
            if (ti[TI_RAND_FRICTION]) DR_PUT(friction, 1, 0, rand_range(tf[TF_FRICTION_LO], tf[TF_FRICTION_SPAN], rng.u(SITE_FRICTION, 0)));
            if (ti[TI_RAND_MASS]) DR_PUT(added_mass, 1, 0, rand_range(tf[TF_MASS_LO], tf[TF_MASS_SPAN], rng.u(SITE_MASS, 0)));
            if (ti[TI_RAND_ARMATURE]) B.joint_armature[env] = rand_range(tf[TF_ARM_LO], tf[TF_ARM_SPAN], rng.u(SITE_ARMATURE, 0));
            if (ti[TI_RAND_JFRICTION]) B.joint_friction[env] = rand_range(tf[TF_JFR_LO], tf[TF_JFR_SPAN], rng.u(SITE_JFRICTION, 0));
            if (ti[TI_RAND_JDAMPING]) B.joint_damping[env] = rand_range(tf[TF_JDA_LO], tf[TF_JDA_SPAN], rng.u(SITE_JDAMPING, 0));
        }
        if (ti[TI_RAND_COM] && lane < 3)
            DR_PUT(com_bias, 3, lane, rand_range(tf[TF_COMX_LO + 2 * lane], tf[TF_COMX_SPAN + 2 * lane], rng.u(SITE_COM, lane)));
        if (ti[TI_RAND_PD] && jl) {
            DR_PUT(kp_scale, A, lane, rand_range(tf[TF_KPS_LO], tf[TF_KPS_SPAN], rng.u(SITE_KP, lane)));
            DR_PUT(kd_scale, A, lane, rand_range(tf[TF_KDS_LO], tf[TF_KDS_SPAN], rng.u(SITE_KD, lane)));
        }
#undef DR_PUT
        if (lane < 3 * F) B.last_feet_vel[env * 3 * F + lane] = 0.f;
        for (int e = lane; e < 48; e += 32) B.contact_warm[env * 48 + e] = 0.f;    // teleported: the contact warm start is stale
        if (ti[TI_CTRL_DELAY]) {                                                     // legged_robot.py:144-148
            const int depth = ti[TI_CTRL_DELAY_HI] + 1, span = ti[TI_CTRL_DELAY_HI] - ti[TI_CTRL_DELAY_LO] + 1;
            for (int e = lane; e < depth * A; e += 32) B.action_queue[(size_t)env * depth * A + e] = 0.f;
            if (lane == 0) B.action_delay[env] = ti[TI_CTRL_DELAY_LO] + min((int)__fmul_rn(rng.u(SITE_CTRL_DELAY, 0), (float)span), span - 1);
        }
        if (ti[TI_BEHAVIOR]) { gtime = 0.f; gphi = 0.f; }                               // go2_wtw.py:139-142
        else if (ti[TI_GAIT]) {                                                     // tron1_pf_ee.py:221-228
            th0 = __fadd_rn(tf[TF_GAIT_THETA_LEFT], rng.u(SITE_GAIT, 0));
            th1 = __fadd_rn(th0, tf[TF_GAIT_THETA_RIGHT]);                              // TF_GAIT_THETA_RIGHT holds right - left
            gtime = __fmul_rn(rng.u(SITE_GAIT, 1), gper);
            gphi = __fdiv_rn(gtime, gper);
        }
        fat = 0.f;
        ep_len = 0; fail_cnt = 0;
    }
    if ((pm & PHASE_RESET) && ti[TI_TERRAIN_CURRICULUM] && !call.force_reset && lane == 0)
        {   // terrain_level means of extras["episode"] (all envs; go2_cts: teacher envs separately, go2_cts.py:93-99): every env
            // contributes every step, so a CTA-synchronous CTA adds them up on chip first (one global reduction per CTA below)
            const bool teacher = env + ti[TI_ENV_OFFSET] < ti[TI_NUM_TEACHER];
            if (cta_sync) { atomicAdd(cta_acc, (float)new_level); if (teacher) atomicAdd(cta_acc + 1, (float)new_level); }
            else { atomicAdd(B.stats + n_sums + 1, (float)new_level); if (teacher) atomicAdd(B.stats + n_sums + 3, (float)new_level); }
        }
    // ---- write back the small per-env state
    if (pm & (PHASE_CALLBACK | PHASE_RESET)) {
        if (lane < 4) B.commands[env * 4 + lane] = lane == 0 ? cmd0 : (lane == 1 ? cmd1 : (lane == 2 ? cmd2 : cmd3));
        if (lane == 0) B.episode_length[env] = ep_len;
    }
    if ((pm & (PHASE_TERMINATION | PHASE_RESET)) && lane == 0) B.fail_buf[env] = fail_cnt;
    if (pm & (PHASE_REWARD | PHASE_RESET)) {
        if (lane < n_sums) B.episode_sums[env * n_sums + lane] = my_sum;
        if (fl) B.feet_air_time[env * F + lane] = fat;
    }
    float clk = 0.f;                 // clock_input[lane]: sin for lanes 0..F-1, cos for lanes F..2F-1 (tron1_pf_ee.py:258-263, go2_wtw.py:251-256)
    if (ti[TI_GAIT] && (pm & PHASE_OBSERVE)) {
        const int fi = lane < F ? lane : lane - F;
        const float thl = fi == 0 ? th0 : (fi == 1 ? th1 : (fi == 2 ? th2 : th3));
        const float arg = __fmul_rn(6.2831853071795862f, __fadd_rn(gphi, thl));
        clk = lane < F ? sinf(arg) : cosf(arg);
        const float clk_sh = __shfl_sync(B200_FULL_MASK, clk, (lane + 32 - B200_GS_CLK) & 31);    // lane B200_GS_CLK + i <- clock of lane i
        if (lane < B200_GAIT_STATE) {
            float v = 0.f;
            if (lane < 4) v = lane == 0 ? th0 : (lane == 1 ? th1 : (lane == 2 ? th2 : th3));
            else if (lane == B200_GS_GT) v = gtime; else if (lane == B200_GS_PHI) v = gphi; else if (lane == B200_GS_PER) v = gper;
            else if (lane == B200_GS_BH) v = bh_t; else if (lane == B200_GS_FC) v = fc_t; else if (lane == B200_GS_PT) v = pt_t;
            else if (lane < B200_GS_CLK + 2 * F) v = clk_sh;
            B.gait_state[env * B200_GAIT_STATE + lane] = v;
        }
    }
    __syncwarp();   // DR parameters written above are re-read below by other lanes

    ENV_SECTION_SYNC();
    // Every reduction extras["episode"] needs has been issued by now.  A CTA-synchronous CTA publishes its on-chip sums and
    // takes its finalize ticket HERE: the fence only has the small state stores above to wait for, and the ticket's round
    // trip runs under the observation phase instead of at the kernel's end, where every CTA of the (single) wave would
    // queue up behind 4 k same-address reductions (ncu: 20 % of the warp time sat at that fence and the barrier after it).
    int ticket = -1;
    if (cta_sync && call.finalize && threadIdx.x == 0) {
        if (cta_acc[0] != 0.f) atomicAdd(B.stats + n_sums + 1, cta_acc[0]);
        if (cta_acc[1] != 0.f) atomicAdd(B.stats + n_sums + 3, cta_acc[1]);
        if (cta_acc[2] != 0.f) atomicAdd(B.stats + n_sums + 2, cta_acc[2]);
        __threadfence();                                        // cumulative: orders the whole CTA's reductions (barrier above)
        ticket = atomicAdd(B.global_flags + 1, 1);
    }
    // ================================================================== compute_observations
    if (pm & PHASE_OBSERVE) {
        const int NO = ti[TI_NUM_OBS];
        float *ob = es + ES_OBS, *nz = es + ES_NOISY;
        if (lane < 3) {
            const float cs = lane < 2 ? tf[TF_OS_LIN_VEL] : tf[TF_OS_ANG_VEL];
            ob[lane] = __fmul_rn(lane == 0 ? cmd0 : (lane == 1 ? cmd1 : cmd2), cs);
            ob[3 + lane] = comp3(grav_obs, lane);
            ob[6 + lane] = __fmul_rn(comp3(ang_b, lane), tf[TF_OS_ANG_VEL]);
        }
        if (jl) {
            ob[9 + lane] = __fmul_rn(__fsub_rn(qj, q0j), tf[TF_OS_DOF_POS]);
            ob[9 + A + lane] = __fmul_rn(qdj, tf[TF_OS_DOF_VEL]);
            ob[9 + 2 * A + lane] = actj;
        }
        __syncwarp();
        const float clipo = tf[TF_CLIP_OBS];
        for (int e = lane; e < ((ti[TI_OBS_KIND] == 4 || ti[TI_OBS_KIND] == 5) ? 0 : NO); e += 32) {
            float v = ob[e];
            if (ti[TI_ADD_NOISE]) {
                const float u = rng.u(SITE_OBS_NOISE, e);
                v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(2.0f, u), 1.0f), tf[TF_NOISE_VEC + e]));
            }
            nz[e] = v;
            const float vc = fminf(fmaxf(v, -clipo), clipo);
            B.obs_buf[env * NO + e] = vc;
            if (call.ro.obs) call.ro.obs[(size_t)env * NO + e] = vc;
        }
        if (ti[TI_OBS_KIND] == 5) {   // go2_wtw.py:53-111: obs_buf = 5 x 61 noisy frames, privileged_obs_buf = 5 x 99 critic frames
            const int SC = ti[TI_SINGLE_CRITIC], NB = 9 + 3 * A;
            float *cr = es + ES_CRIT;
            if (lane < 2 * F) ob[NB + lane] = clk;
            if (lane == 0) { float *d = ob + NB + 2 * F; d[0] = gper; d[1] = bh_t; d[2] = fc_t; d[3] = pt_t; d[4] = th0; d[5] = th1; d[6] = th2; d[7] = th3; }
            __syncwarp();
            for (int e = lane; e < NO; e += 32) {
                float v = ob[e];
                if (ti[TI_ADD_NOISE]) v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(2.0f, rng.u(SITE_OBS_NOISE, e)), 1.0f), tf[TF_NOISE_VEC + e]));
                nz[e] = fminf(fmaxf(v, -clipo), clipo);
                B.obs_buf[env * NO + e] = nz[e];
                cr[e] = fminf(fmaxf(ob[e], -clipo), clipo);
            }
            if (lane < 3) cr[NO + lane] = __fmul_rn(comp3(lin_b, lane), tf[TF_OS_LIN_VEL]);
            if (lane == 0) {
                float *d = cr + NO + 3;
                d[0] = push_vel.x; d[1] = push_vel.y; d[2] = R.added_mass[row]; d[3] = R.friction[row];
                d[4] = R.com_bias[row * 3]; d[5] = R.com_bias[row * 3 + 1]; d[6] = R.com_bias[row * 3 + 2];
            }
            if (jl) { cr[NO + 10 + lane] = R.kp_scale[row * A + lane]; cr[NO + 10 + A + lane] = R.kd_scale[row * A + lane]; }
            if (fl) cr[NO + 10 + 2 * A + lane] = expc_frc;      // exp_C_frc of FL, FR, RL, RR as left by the reward (R13)
            __syncwarp();
            for (int e = lane; e < SC; e += 32) cr[e] = fminf(fmaxf(cr[e], -clipo), clipo);
            __syncwarp();
            ring_put(B.obs_history, env, ti[TI_FRAME_STACK], NO, call.hist_step, nz, lane);
            ring_put(B.critic_obs, env, ti[TI_C_FRAME_STACK], SC, call.hist_step, cr, lane);
        } else if (ti[TI_OBS_KIND] == 4) {   // tron1_pf_ee.py:53-141: features = 10 x 31 noisy frames, labels 17, critic = 10 x 134
            const int SC = ti[TI_SINGLE_CRITIC], NP = ti[TI_NUM_PRIV], NCS = ti[TI_N_CS], NB = NO - 4;   // NB = 9 + 3A
            float *cr = es + ES_CRIT, *pv = es + ES_PRIV;
            const int DRN = 10 + 2 * A;
            if (lane < 2 * F) ob[NB + lane] = clk;
            __syncwarp();
            // the generic loop above already produced nz[0..NB) / obs_buf; redo it over the full 31-wide frame (clock included)
            for (int e = lane; e < NO; e += 32) {
                float v = ob[e];
                if (ti[TI_ADD_NOISE]) v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(2.0f, rng.u(SITE_OBS_NOISE, e)), 1.0f), tf[TF_NOISE_VEC + e]));
                nz[e] = fminf(fmaxf(v, -clipo), clipo);
                B.obs_buf[env * NO + e] = nz[e];
                cr[e] = fminf(fmaxf(ob[e], -clipo), clipo);
            }
            for (int e = lane; e < DRN; e += 32) {
                float v;
                if (e == 0) v = __fsub_rn(R.friction[row], tf[TF_FRICTION_OFFSET]);
                else if (e == 1) v = R.added_mass[row];
                else if (e < 5) v = R.com_bias[row * 3 + e - 2];
                else if (e < 7) v = e == 5 ? push_vel.x : push_vel.y;
                else if (e < 7 + A) v = __fsub_rn(R.kp_scale[row * A + e - 7], tf[TF_KPS_OFFSET]);
                else if (e < 7 + 2 * A) v = __fsub_rn(R.kd_scale[row * A + e - 7 - A], tf[TF_KDS_OFFSET]);
                else v = e == 7 + 2 * A ? B.joint_armature[env] : (e == 8 + 2 * A ? B.joint_friction[env] : B.joint_damping[env]);
                cr[NO + e] = v;
            }
            {   // gait_info: exp_C_frc of the left / right foot as left by the reward (R13)
                const float e0 = __shfl_sync(B200_FULL_MASK, expc_frc, 0), e1 = __shfl_sync(B200_FULL_MASK, expc_frc, 1);
                if (lane == 0) { cr[NO + DRN] = e0; cr[NO + DRN + 1] = e1; }
            }
            int off = NO + DRN + 2;
            for (int e = lane; e < NCS; e += 32) { cr[off + e] = es[ES_LCS + e]; pv[3 + e] = es[ES_LCS + e]; }
            off += NCS;
            for (int pt = lane; pt < P; pt += 32) {
                const float d = __fsub_rn(__fsub_rn(bp.z, tf[TF_HEIGHT_OBS_OFFSET]), es[ES_MH + pt]);
                cr[off + pt] = __fmul_rn(fminf(fmaxf(d, -1.0f), 1.0f), tf[TF_OS_HEIGHT]);
            }
            off += P;
            for (int e = lane; e < 3 * F; e += 32) { cr[off + e] = es[ES_NV + e]; pv[3 + NCS + F + e] = es[ES_NV + e]; }
            off += 3 * F;
            for (int e = lane; e < 9 * F; e += 32) {
                const float fz = R.feet_pos[(row * F + e / 9) * 3 + 2];
                cr[off + e] = fminf(fmaxf(__fsub_rn(fz, es[ES_HAF + e]), -1.0f), 1.0f);
            }
            if (lane < 3) pv[lane] = __fmul_rn(comp3(lin_b, lane), tf[TF_OS_LIN_VEL]);
            if (fl) pv[3 + NCS + lane] = fminf(fmaxf(__fsub_rn(__fsub_rn(fpos.z, hmax), tf[TF_FOOT_HEIGHT_OFFSET]), -1.0f), 1.0f);
            __syncwarp();
            for (int e = lane; e < NP; e += 32) {                                                   // estimator labels (unclipped)
                B.privileged_obs_buf[env * NP + e] = pv[e];
                if (call.ro.privileged_obs) call.ro.privileged_obs[(size_t)env * NP + e] = pv[e];
            }
            for (int e = lane; e < SC; e += 32) cr[e] = fminf(fmaxf(cr[e], -clipo), clipo);
            __syncwarp();
            ring_put(B.obs_history, env, ti[TI_FRAME_STACK], NO, call.hist_step, nz, lane);
            ring_put(B.critic_obs, env, ti[TI_C_FRAME_STACK], SC, call.hist_step, cr, lane);
        } else if (ti[TI_OBS_KIND] == 3) {   // tron1_pf.py:15-70: obs_buf = stack of noisy frames, privileged_obs_buf = stack of critic frames
            const int SC = ti[TI_SINGLE_CRITIC];
            float *cr = es + ES_CRIT;
            const bool cleared = (pm & PHASE_RESET) && reset;
            if (lane < 3) cr[lane] = fminf(fmaxf(__fmul_rn(comp3(lin_b, lane), tf[TF_OS_LIN_VEL]), -clipo), clipo);
            for (int e = lane; e < NO; e += 32) { cr[3 + e] = fminf(fmaxf(ob[e], -clipo), clipo); nz[e] = fminf(fmaxf(nz[e], -clipo), clipo); }
            if (jl) cr[3 + NO + lane] = cleared ? 0.f : lastj;
            if (lane == 0) {
                float *d = cr + 3 + NO + A;
                d[0] = __fsub_rn(R.friction[row], tf[TF_FRICTION_OFFSET]); d[1] = R.added_mass[row];
                d[2] = R.com_bias[row * 3]; d[3] = R.com_bias[row * 3 + 1]; d[4] = R.com_bias[row * 3 + 2];
                d[5] = push_vel.x; d[6] = push_vel.y;
            }
            if (fl) cr[3 + NO + A + 7 + lane] = fat;
            __syncwarp();
            ring_put(B.obs_history, env, ti[TI_FRAME_STACK], NO, call.hist_step, nz, lane);
            ring_put(B.critic_obs, env, ti[TI_C_FRAME_STACK], SC, call.hist_step, cr, lane);
        } else if (ti[TI_OBS_KIND] >= 1) {   // go2_ts (1) / go2_cat (2) / go2_cts (6) / go2_ee (7) / go2_dreamwaq (8)
            const int kind = ti[TI_OBS_KIND];
            const bool cat = kind == 2;      // go2_cat.py:19-99: 3 more DR entries, no base_lin_vel, raw feet heights
            const bool cts = kind == 6;      // go2_cts.py:36-91: go2_ts with raw feet heights in the privileged obs
            const bool ee = kind == 7;       // go2_ee.py:10-73: critic without base_lin_vel, estimator labels, clipped stacks
            const bool dwq = kind == 8;      // go2_dreamwaq.py:7-90: base_lin_vel leads the critic frame, labels, next state
            const bool labels = ee || dwq;
            const int SC = ti[TI_SINGLE_CRITIC], NP = ti[TI_NUM_PRIV], NCS = ti[TI_CONTACT_STATES] ? ti[TI_N_CS] : 0;
            float *cr = es + ES_CRIT, *pv = es + ES_PRIV;
            const int DRN = (cat ? 10 : 7) + 2 * A, LIN = (cat || ee) ? 0 : 3;
            const int c_ob = dwq ? 3 : 0, c_lin = dwq ? 0 : NO + DRN;     // critic frame = [ob, dr, lin | lin, ob, dr], lcs, heights
            for (int e = lane; e < NO; e += 32) cr[c_ob + e] = ob[e];
            // domain_randomization_info (go2_ts.py:16-28)
            for (int e = lane; e < DRN; e += 32) {
                float v;
                if (e == 0) v = __fsub_rn(R.friction[row], tf[TF_FRICTION_OFFSET]);
                else if (e == 1) v = R.added_mass[row];
                else if (e < 5) v = R.com_bias[row * 3 + e - 2];
                else if (e < 7) v = e == 5 ? push_vel.x : push_vel.y;
                else if (e < 7 + A) v = __fsub_rn(R.kp_scale[row * A + e - 7], tf[TF_KPS_OFFSET]);
                else if (e < 7 + 2 * A) v = __fsub_rn(R.kd_scale[row * A + e - 7 - A], tf[TF_KDS_OFFSET]);
                else v = e == 7 + 2 * A ? B.joint_armature[env] : (e == 8 + 2 * A ? B.joint_friction[env] : B.joint_damping[env]);
                cr[c_ob + NO + e] = v;
                if (!labels) pv[e] = v;
            }
            if (lane < 3) {
                const float v = __fmul_rn(comp3(lin_b, lane), tf[TF_OS_LIN_VEL]);
                if (LIN) cr[c_lin + lane] = v;
                if (labels) pv[lane] = dwq ? __fmul_rn(v, 0.5f) : v;                  // go2_ee.py:67 / go2_dreamwaq.py:84
                else if (LIN) pv[DRN + 12 * F + lane] = v;
            }
            for (int e = lane; e < NCS; e += 32) {
                const float v = es[ES_LCS + e];
                cr[NO + DRN + LIN + e] = v;
                if (labels) pv[3 + e] = v; else pv[DRN + 12 * F + LIN + e] = v;
            }
            if (ti[TI_MEASURE_HEIGHTS]) {
                for (int pt = lane; pt < P; pt += 32) {
                    const float d = __fsub_rn(__fsub_rn(bp.z, tf[TF_HEIGHT_OBS_OFFSET]), es[ES_MH + pt]);
                    cr[NO + DRN + LIN + NCS + pt] = __fmul_rn(fminf(fmaxf(d, -1.0f), 1.0f), tf[TF_OS_HEIGHT]);
                }
            }
            if (labels) {                                                            // foot clearance above the mean terrain height
                if (fl) pv[3 + NCS + lane] = fminf(fmaxf(__fsub_rn(__fsub_rn(fpos.z, hmean), tf[TF_FOOT_HEIGHT_OFFSET]), -1.0f), 1.0f);
            } else {
                for (int e = lane; e < 9 * F; e += 32) {
                    const float fz = R.feet_pos[(row * F + e / 9) * 3 + 2];
                    pv[DRN + e] = (cat || cts) ? es[ES_HAF + e] : fminf(fmaxf(__fsub_rn(fz, es[ES_HAF + e]), -1.0f), 1.0f);
                }
                for (int e = lane; e < 3 * F; e += 32) pv[DRN + 9 * F + e] = es[ES_NV + e];
            }
            if (dwq) {                                                               // decoder target, go2_dreamwaq.py:72-80
                for (int e = lane; e < 9 + 2 * A; e += 32) { B.next_state_buf[env * NO + e] = ob[e]; if (call.ro.next_state) call.ro.next_state[(size_t)env * NO + e] = ob[e]; }
                if (jl) {
                    const float vn = __fmul_rn(actj, tf[TF_ACTION_SCALE]);
                    B.next_state_buf[env * NO + 9 + 2 * A + lane] = vn;
                    if (call.ro.next_state) call.ro.next_state[(size_t)env * NO + 9 + 2 * A + lane] = vn;
                }
            }
            __syncwarp();
            // estimator labels are handed out unclipped (legged_robot_ee.py:56-73, legged_robot_dreamwaq.py:63-79)
            for (int e = lane; e < NP; e += 32) {
                const float vp = labels ? pv[e] : fminf(fmaxf(pv[e], -clipo), clipo);
                B.privileged_obs_buf[env * NP + e] = vp;
                if (call.ro.privileged_obs) call.ro.privileged_obs[(size_t)env * NP + e] = vp;
            }
            if (labels) {       // these steps return the clipped stacks: clip the frame going in (ee: both stacks, dreamwaq: critic only)
                for (int e = lane; e < SC; e += 32) cr[e] = fminf(fmaxf(cr[e], -clipo), clipo);
                if (ee) for (int e = lane; e < NO; e += 32) nz[e] = fminf(fmaxf(nz[e], -clipo), clipo);
                __syncwarp();
            }
            // history stacks: the new frame goes into both of its ring slots (legged_robot_ts.py:29-47)
            ring_put(B.obs_history, env, ti[TI_FRAME_STACK], NO, call.hist_step, nz, lane);
            ring_put(B.critic_obs, env, ti[TI_C_FRAME_STACK], SC, call.hist_step, cr, lane);
        }
    }
    // Tasks that shift the history again at the end of post_physics_step (go2_cat.py:127-130): the dynamics
    // kernel shifts once more at the start of the next step, so llast == last at reward time.
    if (ti[TI_DOUBLE_SHIFT] && (pm & PHASE_OBSERVE)) {
        const bool was_reset = (pm & PHASE_RESET) && reset;
        if (jl) {
            B.llast_actions[env * A + lane] = was_reset ? 0.f : lastj;
            B.last_actions[env * A + lane] = actj;
            B.last_dof_vel[env * A + lane] = qdj;
        }
        if (fl) { float *lv = B.last_feet_vel + (env * F + lane) * 3; lv[0] = fvel.x; lv[1] = fvel.y; lv[2] = fvel.z; }
    }
    return ticket;
}

// Bug-compatible mode only (TaskSpec.reproduce_r18; never on the default path): the reference indexes per-env
// gait tensors with the flattened nonzero() of [N,1] masks, so row 0 follows "any env" for the gait-clock wrap and for the
// swing / stance indicator of each foot.  One thread per env evaluates exactly what env_post_step_warp is about to use --
// the gait state after this step's behaviour resampling in the callback (go2_wtw.py:258-263) -- and ORs the bits into
// global_flags[0] (bit 8: wrap, 9..12: swing per foot, 13..16: stance per foot); launched right before the env kernel.
#ifdef B200_WARP_EMU
#define R18_FLAGS_BLOCK 32          // the test emulator runs one warp per block
#else
#define R18_FLAGS_BLOCK 128
#endif
#ifdef B200_ENV_DEFINE_KERNELS
__global__ void r18_flags_kernel(const TaskDev T, const B200Buffers B, const EnvCall call) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    int bits = 0;
    if (env < T.i[TI_NUM_ENVS]) {
        const float *tf = T.f;
        const float *gs = B.gait_state + (size_t)env * B200_GAIT_STATE;
        float th[4] = {gs[B200_GS_TH], gs[B200_GS_TH + 1], gs[B200_GS_TH + 2], gs[B200_GS_TH + 3]};
        float gper = T.i[TI_BEHAVIOR] ? gs[B200_GS_PER] : tf[TF_GAIT_PERIOD];
        const float dt = tf[TF_POLICY_DT];
        if (T.i[TI_BEHAVIOR] && (call.phase_mask & PHASE_CALLBACK) && (B.episode_length[env] + 1) % T.i[TI_BEHAVIOR_INTERVAL] == 0) {
            EnvRng rng; rng.k0 = (uint32_t)T.i[TI_SEED_LO]; rng.k1 = (uint32_t)T.i[TI_SEED_HI]; rng.env = (uint32_t)(env + T.i[TI_ENV_OFFSET]); rng.step = call.step;
            gper = rand_range(call.beh[0], call.beh[1], rng.u(SITE_BEHAVIOR, 0));
            for (int k = 0; k < 4; k++) th[k] = tf[TF_GAIT_THETA + call.gait_cb * B200_MAX_FEET + k];
        }
        if (__fadd_rn(gs[B200_GS_GT], dt) >= __fsub_rn(gper, __fmul_rn(dt, 0.5f))) bits |= 1;
        for (int f = 0; f < T.i[TI_F]; f++) {
            const float ph = __fmul_rn(fmodf(__fadd_rn(gs[B200_GS_PHI], th[f]), 1.0f), 6.2831853071795862f);
            if (ph >= 0.f && ph < tf[TF_GAIT_B_SWING]) bits |= 2 << f;
            if (ph >= tf[TF_GAIT_B_SWING] && ph < 6.2831853071795862f) bits |= 32 << f;
        }
    }
    const unsigned any = __ballot_sync(B200_FULL_MASK, bits != 0);
    if (any) {                                                  // OR over the warp (bug-compat path: clarity over speed)
        for (int o = 16; o > 0; o >>= 1) bits |= __shfl_xor_sync(B200_FULL_MASK, bits, o);
        if ((threadIdx.x & 31) == 0) atomicOr(B.global_flags, bits << 8);
    }
}
#else
__global__ void r18_flags_kernel(const TaskDev T, const B200Buffers B, const EnvCall call);
#endif   // B200_ENV_DEFINE_KERNELS

// extras["episode"] (legged_robot.py:127-141): means over the envs that reset this step, from the reductions the env
// kernel left in stats[0..n_sums] (+ reset count, + sum of terrain levels); written to slot (step % ENV_STATS_RING) of
// the ring that follows the work area so that the host can hand out per-step values without any further launch.
// Run by the last CTA of env_post_step_kernel (thread i < n_sums + B200_STATS_EXTRA writes entry i).
__device__ __forceinline__ void stats_finalize(float *stats, int n_sums, const EnvCall &call, int i) {
    float *ring = stats + (2 * n_sums + 4) + call.stats_slot * (n_sums + B200_STATS_EXTRA);
    const float cnt = __ldcg(stats + n_sums);
    // no env reset in this step: the reference leaves the previous extras["episode"]["rew_*"] in place (reset_idx returns
    // early, legged_robot.py:105-106), so the slot carries the previous step's means forward
    const float *prev = stats + (2 * n_sums + 4) + ((call.stats_slot + ENV_STATS_RING - 1) % ENV_STATS_RING) * (n_sums + B200_STATS_EXTRA);
    if (i < n_sums) ring[i] = cnt > 0.f ? __ldcg(stats + i) / cnt * call.inv_episode_length_s : __ldcg(prev + i);
    if (i == n_sums) ring[i] = __ldcg(stats + n_sums + 1) * call.inv_num_envs;            // mean terrain level
    if (i == n_sums + 1) ring[i] = call.inv_teacher > 0.f ? __ldcg(stats + n_sums + 3) * call.inv_teacher     // go2_cts: teacher terrain level
                                                          : __ldcg(stats + n_sums + 2) * call.inv_num_envs;   // mean CaT termination probability
    if (i == n_sums + 2) ring[i] = (__ldcg(stats + n_sums + 1) - __ldcg(stats + n_sums + 3)) * call.inv_student;  // go2_cts: student terrain level
    if (i == n_sums + 3) ring[i] = cnt;                                                   // envs that reset this step (0: the rew_* entries are carried over)
}

// A CTA that took its ticket early (env_post_step_warp): its first warp finalises if the ticket was the last one.  No CTA
// barrier, no fence unless last -- the other warps just leave.
// rew | reset | time_out of every env into the caller's pinned host slab (EnvCall::host_slab): run by the whole CTA that
// took the last ticket -- every other CTA wrote its rows before its fence and ticket -- with all loads (L2, not this SM's
// L1) in flight before the first store.
static __device__ __noinline__ void env_host_slab_copy(const uint4 *src, uint4 *dst, int vecs) {    // out of line: run by one CTA per launch,
    constexpr int BATCH = 6;                                                                    // must not shape the kernel's registers
    for (int base = 0; base < vecs; base += BATCH * (int)blockDim.x) {
        uint4 v[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) v[k] = __ldcg(src + min(base + k * (int)blockDim.x + (int)threadIdx.x, vecs - 1));
#pragma unroll
        for (int k = 0; k < BATCH; k++) { const int i = base + k * (int)blockDim.x + (int)threadIdx.x; if (i < vecs) dst[i] = v[k]; }
    }
}

__device__ __forceinline__ void env_finalize_ticket(const B200Buffers &B, int n_sums, const EnvCall &call, int ticket, int *last_w) {
    if (call.host_slab) {                // uniform over the grid: one more CTA barrier, only on the host-buffer path
        if (threadIdx.x == 0) *last_w = ticket == (int)gridDim.x - 1;
        __syncthreads();
        if (!*last_w) return;
        __threadfence();
        env_host_slab_copy((const uint4 *)B.rew_buf, call.host_slab, call.host_slab_vecs);
        if (threadIdx.x >= 32) return;
    } else {
        if (threadIdx.x >= 32) return;
        const bool last = __shfl_sync(B200_FULL_MASK, ticket, 0) == (int)gridDim.x - 1;
        if (!last) return;
        __threadfence();
    }
    for (int i = (int)threadIdx.x; i < n_sums + B200_STATS_EXTRA; i += 32) stats_finalize(B.stats, n_sums, call, i);
    if (threadIdx.x == 0) B.global_flags[1] = 0;
}

// The CTA that takes the last ticket sees every CTA's reductions (fence + atomic) and finalises them.  Called by every
// thread of every CTA; `last` is one shared-memory word nobody else uses at this point.
__device__ __forceinline__ void env_finalize_cta(const B200Buffers &B, int n_sums, const EnvCall &call, int *last) {
    __syncthreads();           // the CTA's reductions are ordered before thread 0 ...
    if (threadIdx.x == 0) {    // ... whose fence (cumulative) publishes them device-wide before it takes the ticket
        __threadfence();
        *last = atomicAdd(B.global_flags + 1, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (*last) {
        __threadfence();
        if (call.host_slab) env_host_slab_copy((const uint4 *)B.rew_buf, call.host_slab, call.host_slab_vecs);
        stats_finalize(B.stats, n_sums, call, (int)threadIdx.x);
        if (threadIdx.x == 0) B.global_flags[1] = 0;
    }
}

template <class S>
__device__ __forceinline__ void env_post_step_body_impl(const TaskDev &T, const B200Buffers &B, const TerrainDev &tr, const EnvCall &call, const EnvStageTab &tab) {
    extern __shared__ float smem[];
    const TiView<S> ti{T.i};
    constexpr int nwarps = ENV_WARPS_PER_BLOCK;            // the launch geometry is fixed: slab offsets fold to constants
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // (pinning them like the dynamics kernel does costs 1 us here)
    const int env0 = blockIdx.x * nwarps, env = env0 + warp;
    const int N = ti[TI_NUM_ENVS];
    uint64_t *bar = (uint64_t *)smem;
    float *es = smem + ENV_HDR_WORDS, *cta_acc = smem + ENV_HDR_ACC;
    char *inslab = (char *)(es + nwarps * ES_TOTAL);
    const bool full = !call.force_reset && (call.phase_mask & PHASE_ALL) == PHASE_ALL;
    const bool staged = full && tab.ok && env0 + nwarps <= N;
    if (staged) {
        if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_expect_tx(bar, tab.in_bytes); cta_acc[0] = 0.f; cta_acc[1] = 0.f; cta_acc[2] = 0.f; }
        __syncthreads();                                   // barrier armed before any copy can complete on it
        grid_dependency_wait();                            // launched under the dynamics kernel (b200_env_step): its grid is complete from here on
        // thread t issues the bulk copy of table entry t
        for (int t = (int)threadIdx.x; t < tab.n; t += (int)blockDim.x)
            if (tab.off[t] != ENV_NOT_STAGED)
                tma_load_1d(inslab + tab.off[t], tab.src[t] + (size_t)env0 * tab.row_bytes[t], (uint32_t)nwarps * tab.row_bytes[t], bar);
        // every thread builds the same read view: staged tensors lead into the slab (offsets are compile-time constants in
        // the specialised instantiations), the few that could not be staged are biased to the CTA's first env; row = warp
        EnvInputs R;
        int t = 0;
#define X_VIEW(field, type, k) { const uint32_t o_ = S::fixed ? env_stage_layout(t, ti[TI_A], ti[TI_F], ti[TI_L], ti[TI_N_SUMS], nwarps) : tab.off[t]; \
        R.field = o_ != ENV_NOT_STAGED ? (const type *)(inslab + o_) : B.field + (size_t)env0 * (k); t++; }
        ENV_STAGED_INPUTS(X_VIEW, ti[TI_A], ti[TI_F], ti[TI_L], ti[TI_N_SUMS])
#undef X_VIEW
#ifndef ENV_NO_SECTION_SYNC
        const int ticket = env_post_step_warp<S>(T, B, R, warp, tr, call, es + warp * ES_TOTAL, env, lane, true, true, bar, cta_acc);
        if (call.finalize) env_finalize_ticket(B, ti[TI_N_SUMS], call, ticket, (int *)(smem + 2));
        return;
#else
        env_post_step_warp<S>(T, B, R, warp, tr, call, es + warp * ES_TOTAL, env, lane, true, false, bar, cta_acc);
#endif
    } else {
        grid_dependency_wait();                            // every thread, also the idle warps of a ragged last CTA
        if (env < N) {
            EnvInputs R;
#define X_VIEW(field, type, k) R.field = B.field;
            ENV_STAGED_INPUTS(X_VIEW, 0, 0, 0, 0)
#undef X_VIEW
            env_post_step_warp<S>(T, B, R, env, tr, call, es + warp * ES_TOTAL, env, lane, false, false, bar, cta_acc);
        }
    }
    if (call.finalize) env_finalize_cta(B, ti[TI_N_SUMS], call, (int *)(smem + 2));
}

// DEV: a device-stepped launch.  The per-step scalars come from the step state in device memory (EnvCall::dstate) -- four
// uniform 16-byte loads per thread -- into a private copy of the call; the host-stepped instantiations read the kernel
// parameter where it lies (constant bank operands, no registers), which is why the two are separate instantiations.
template <class S, bool DEV>
__device__ __forceinline__ void env_post_step_body(const TaskDev &T, const B200Buffers &B, const TerrainDev &tr, const EnvCall &call_in, const EnvStageTab &tab) {
    if constexpr (DEV) {
        EnvCall call = call_in;
        const int4 *ss = (const int4 *)call_in.dstate;
        const int4 a = __ldg(ss), b = __ldg(ss + 1), c = __ldg(ss + 2), d = __ldg(ss + 3);
        call.step = (uint32_t)a.x; call.hist_step = (uint32_t)a.y - 1u;     // [1] already counts this call's frame
        call.vx_lo = __int_as_float(a.z); call.vx_span = __int_as_float(a.w);
        call.beh[0] = __int_as_float(b.x); call.beh[1] = __int_as_float(b.y); call.beh[2] = __int_as_float(b.z); call.beh[3] = __int_as_float(b.w);
        call.beh[4] = __int_as_float(c.x); call.beh[5] = __int_as_float(c.y); call.beh[6] = __int_as_float(c.z); call.beh[7] = __int_as_float(c.w);
        call.sit_pose = d.x; call.gait_cb = d.y; call.gait_reset = d.z;
        call.stats_slot = (int)(call.step % (uint32_t)ENV_STATS_RING);
        env_post_step_body_impl<S>(T, B, tr, call, tab);
    } else {
        env_post_step_body_impl<S>(T, B, tr, call_in, tab);
    }
}

// generic instantiation: every descriptor int is read at run time (any task configuration)
template <bool DEV>
__global__ void B200_LAUNCH_BOUNDS(ENV_WARPS_PER_BLOCK * 32, ENV_MIN_BLOCKS)
env_post_step_kernel(const TaskDev T, const B200Buffers B, const TerrainDev tr, const EnvCall call, const EnvStageTab tab) {
    env_post_step_body<EnvSpecGeneric, DEV>(T, B, tr, call, tab);
}

// one instantiation per built-in preset (env_presets.inc); selected by b200_create when the descriptor matches the table
template <int P, bool DEV>
__global__ void B200_LAUNCH_BOUNDS(ENV_WARPS_PER_BLOCK * 32, ENV_MIN_BLOCKS)
env_post_step_kernel_preset(const TaskDev T, const B200Buffers B, const TerrainDev tr, const EnvCall call, const EnvStageTab tab) {
    env_post_step_body<EnvPresetSpec<P>, DEV>(T, B, tr, call, tab);
}

typedef void (*EnvKernelFn)(const TaskDev, const B200Buffers, const TerrainDev, const EnvCall, const EnvStageTab);
// The instantiations live in translation units of their own, compiled beside b200_step.cu and with -fmad=false: the env half
// restates torch fp32 eager arithmetic, where every operation rounds on its own -- left to ptxas, which fuses FMUL + FADD
// where its schedule likes it, two instantiations of the same source (preset / generic, host- / device-stepped) differ in the
// last bit of a reward.  (fmaf() calls stay fused; the dynamics kernel keeps contraction.)
EnvKernelFn env_kernel_fn(int preset);        // csrc/env_kernels_host.cu (+ r18_flags_kernel, step_advance_kernel)
EnvKernelFn env_kernel_dev_fn(int preset);    // csrc/env_kernels_dev.cu

// index of the preset whose compile-time ints all equal this descriptor's, or -1 (-> generic kernel)
inline int env_match_preset(const int *ti) {
    for (int p = 0; p < ENV_NUM_PRESETS; p++) {
        bool same = true;
        for (int k = 0; k < TI_COUNT && same; k++) same = env_ti_is_runtime(k) || kEnvPresetTI[p][k] == ti[k];
        if (same) return p;
    }
    return -1;
}

