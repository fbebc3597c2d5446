// dynamics_kernel.cuh -- decimated PD torque loop + articulated rigid-body substeps, one env per warp.
//
// Replaces, for one policy step (reference paths relative to the LeggedGym-Ex root):
//   LeggedRobot._pre_sim_step              legged_gym/envs/base/legged_robot.py:230-252   (clip, action history)
//   GenesisSimulator.step                  legged_gym/simulator/genesis_simulator.py:20-33 (last_* copy, decimation loop)
//   GenesisSimulator._compute_torques      legged_gym/simulator/genesis_simulator.py:630-642
//   scene.step() of the third-party engine legged_gym/simulator/genesis_simulator.py:29    (formulation: DESIGN.md)
//   get_links_pos/vel/net_contact_force    legged_gym/simulator/genesis_simulator.py:49-51
//
// Mapping to the hardware.  A warp owns one env for all `decimation` substeps; the env's state is parked in the warp's
// shared-memory scratch between phases, HBM is touched once on entry and once on exit (~1.3 KB per env).  The robot is a
// star: a floating base plus C serial chains of 3 revolute joints, so
//   * the chains are walked body by body for what must follow the chain (frames, velocities, velocity-product
//     accelerations; every lane walks one, lanes >= C shadow the last) with all spatial quantities expressed about the base
//     origin in world axes -- composite inertias and forces then add without any frame transforms -- and the body-local
//     half of the RNE (world inertia, momenta, bias wrench) is done by ONE LANE PER BODY, all bodies at once;
//   * the mass matrix is an arrowhead [diag(D_c) B; B^T M_bb]: each chain lane inverts its own 3x3 block, the
//     6x6 Schur complement is summed through shared memory and factorised redundantly by every lane (no sync needed);
//   * collision spheres are tested two per lane against the L2-resident int16 heightfield, the active set is
//     compacted with ballots;
//   * each of the <=32 constraint rows (3 per contact + joint limits + frictionloss) lives in one lane: the lane forms
//     rb = Jb - G^T Jl, yb = S^-1 rb and D^-1 Jl for its row (never the chain parts of M^-1 J^T for the other chains) and
//     its column of A = J M^-1 J^T; the projected BLOCK Gauss-Seidel sweep solves a contact's three rows together from three
//     shuffled residuals (friction-cone projection per contact), unrolled over the <= 8 contacts;
//   * per-warp scratch (frames, M^-1 factors, J rows, A) is staged in shared memory, the robot model once per CTA;
//     CTAs of 14 warps, two per SM: 28 envs per SM, 4096 envs in one wave.
// Redundancy ACROSS lanes is free on a SIMT machine; what costs is repetition ALONG the instruction stream.
// The kernel is latency/issue bound, not HBM bound; tensor cores do not apply.
#pragma once
#include "task_dev.cuh"

#ifndef DYN_WARPS_PER_BLOCK
#define DYN_WARPS_PER_BLOCK 14     // x DYN_MIN_BLOCKS = 2 CTAs per SM: 28 envs per SM, 4096 envs in one wave of 293 CTAs.  Round 1's kernel
#endif                             // preferred 7 x 4; the current one measures 14 x 2 ahead at every env count (profiles/r2ak_sweep_dynamics_cta_shape.log:
                                   // dynamics kernel 0.1454 / 0.1414 / 0.1474 ms for 7x4 / 14x2 / 28x1 at 4096 envs, 1.95 / 1.84 / 1.95 ms at 65 536,
                                   // where CTAs in different phases compete for an SM's instruction cache)
// The substep body is ~7k SASS instructions of mostly straight-line code (~110 KB), larger than the 32 KB L1.5
// instruction cache: warps that drift apart each stream it from L2 and the kernel becomes instruction-fetch bound
// (ncu: stall_no_instruction on top, profiles/r1b).  A CTA barrier at the top of a substep and after its
// contact solve keeps the warps of a CTA in the same code region so one fetch serves all of them.
#ifndef DYN_NO_PHASE_SYNC
#define PHASE_SYNC() __syncthreads()
#define DYN_PHASE_SYNCS_PER_SUBSTEP 5
#else
#define PHASE_SYNC()
#endif
#ifndef DYN_SYNC_MASK
#define DYN_SYNC_MASK 16          // which of the five phase barriers are kept: A top of the substep, B after FK, C before collision, D before
                                  // the rows, E after the solve.  Round 1: A + E did all the good (profiles/r1B_sweep_sync_mask.log: none 0.313,
                                  // A 0.284, A+E 0.281, all five 0.284 ms per step); with the current kernel E alone is ahead
                                  // (profiles/r2ao_sweep_sync_mask.log: dynamics kernel 0.1414 / 0.1393 / 0.1423 ms for A+E / E / all five at
                                  // 4096 envs, 1.837 / 1.806 / 1.867 ms at 65 536): the integration between E and the next substep's top is short
                                  // and uniform, so A re-aligns warps that E has just aligned; no barrier at all 0.1617 ms, only C 0.1516, only D 0.1531
#endif
#define PHASE_SYNC_A() do { if (DYN_SYNC_MASK & 1) PHASE_SYNC(); } while (0)
#define PHASE_SYNC_B() do { if (DYN_SYNC_MASK & 2) PHASE_SYNC(); } while (0)
#define PHASE_SYNC_C() do { if (DYN_SYNC_MASK & 4) PHASE_SYNC(); } while (0)
#define PHASE_SYNC_D() do { if (DYN_SYNC_MASK & 8) PHASE_SYNC(); } while (0)
#define PHASE_SYNC_E() do { if (DYN_SYNC_MASK & 16) PHASE_SYNC(); } while (0)
#ifndef DYN_PGS_AREG
#define DYN_PGS_AREG 4            // contacts whose entries of A a lane keeps in registers during the solve (0: all from shared memory)
#endif
#define DYN_MODE_SIM_ONLY 1        // b200_simulator_step (plugin mode): no pre-step book-keeping
#define DYN_MODE_LATE_ACTIONS 2    // the action buffer is filled by fetch_actions_kernel, this launch is its programmatic dependent
#ifndef DYN_MIN_BLOCKS
#define DYN_MIN_BLOCKS 2      // resident CTAs per SM the register allocator must allow: 2 x 14 warps = 28 envs per SM at 72 regs/thread
#endif

// ---- per-warp shared scratch (floats) ----
// The frames / joint axes / body velocities / contact list / aux rows are consumed by the time the constraint rows are
// built; the Delassus matrix A is written after that point and is dead before the next substep's kinematics rewrites
// them, so A shares their storage.  7 KB per env lets 28 envs (7 CTAs of 4 warps) live on one SM: at 4096 envs per GPU
// the whole batch is resident in a single wave.
#define WS_FR 0                         // frames [13][12]  R(9) o(3)
#define WS_AX (WS_FR + 13 * 12)         // joint axes [12][3]
#define WS_VEL (WS_AX + 36)             // body spatial velocity [13][6] (w, v)
#define WS_CT (WS_VEL + 78)             // contacts [8][12]
#define WS_AUX (WS_CT + 96)             // aux rows [8][4]: joint, sign, pos, bound
#define WS_AM 0                         // A [32][33], aliases everything above (398 floats)
#define WS_MI (WS_AM + 32 * 33)         // per chain: Dinv(6) G(18) ; then Sinv(21)
#define WS_NU (WS_MI + 4 * 24 + 24)     // generalized velocity [18]
#define WS_AF (WS_NU + 18)              // smooth acceleration [18]
#define WS_JR (WS_AF + 18)              // J rows [32][10]: Jb(6) Jl(3) chain
#define WS_FV (WS_JR + 320)             // row force vectors [32][4]
#define WS_LF (WS_FV + 128)             // link forces [17*3]
#define WS_WARM (WS_LF + 52)            // PGS warm start [48]: 8 x (sphere id + 1, f_n, f_t1, f_t2), 8 x (aux code + 1, f)
#define WS_PD (WS_WARM + 48)             // per chain lane [4][16]: tgt(3) kp(3) kd(3) arm+h*dmp(3) dmp(3) frictionloss(1) -- constant over the substeps,
                                         // kept here rather than in 15 registers that are live across the whole substep
// The env's state between the phases of a substep.  It used to live in ~30 registers across the row assembly and the contact
// solve, which the 72-register cap turned into local-memory spills -- and with 186 KB of stack per SM against ~36 KB of L1
// those reloads came from L2 (ncu: 30 % hit rate, long-scoreboard 13.6 % of the warp time).  Parked here, each phase reloads
// what it needs with broadcast LDS.128 at shared-memory latency.
#define WS_ST (WS_PD + 64)               // [20]: p(3) mass_add | Q wxyz(4) | vb(3) mu | wb(3) - | com_shift(3) -
#define WS_QS (WS_ST + 20)               // per chain lane [4][12]: q(3) - | qd(3) - | tau(3) -
#define WS_TOTAL (WS_QS + 48)
static_assert((WS_ST % 4) == 0 && (WS_QS % 4) == 0 && (WS_TOTAL % 4) == 0, "16-byte aligned vector loads");
static_assert(WS_AUX + 32 <= WS_MI, "aliased inputs must fit under the A matrix");

#define MS_BODY 0
#define MS_LINK (MS_BODY + B200_MAX_BODIES * B200_BODY_STRIDE)
#define MS_SPH (MS_LINK + B200_MAX_LINKS * 3)
#define MS_INT (MS_SPH + B200_MAX_SPHERES * 4)      // ints: link_body[17], sph_body[64], sph_link[64]
#define MS_TOTAL (MS_INT + B200_MAX_LINKS + 2 * B200_MAX_SPHERES)

__host__ __device__ inline int dyn_smem_bytes(int warps) { return (MS_TOTAL + warps * WS_TOTAL) * 4; }

#ifdef DYN_TIMING     // diagnosis build (tools/probe_dyn_timing.py): per-env start / end time, sweeps, rows
__device__ unsigned long long g_dyn_timing[65536 * 4];
__device__ __forceinline__ unsigned long long dyn_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
struct SIn { float m; f3 h; float I[6]; };  // spatial inertia about O: I = xx yy zz xy xz yz

__device__ __forceinline__ void si_apply(const SIn &s, f3 w, f3 v, f3 &L, f3 &P) {
    P = v * s.m + cross3(w, s.h);
    L = mk3(s.I[0] * w.x + s.I[3] * w.y + s.I[4] * w.z, s.I[3] * w.x + s.I[1] * w.y + s.I[5] * w.z,
            s.I[4] * w.x + s.I[5] * w.y + s.I[2] * w.z) + cross3(s.h, v);
}
__device__ __forceinline__ void si_acc(SIn &a, const SIn &b) {
    a.m += b.m; a.h = a.h + b.h;
#pragma unroll
    for (int k = 0; k < 6; k++) a.I[k] += b.I[k];
}
// body inertia (mass m, com c rel. O, local inertia Il rotated by R) about O in world axes
__device__ __forceinline__ SIn si_body(float m, f3 c, const float *Il, const m33 &R) {
    SIn s; s.m = m; s.h = c * m;
    // Iw = R Il R^T
    float t[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        t[3 * i + 0] = R.m[3 * i] * Il[0] + R.m[3 * i + 1] * Il[3] + R.m[3 * i + 2] * Il[4];
        t[3 * i + 1] = R.m[3 * i] * Il[3] + R.m[3 * i + 1] * Il[1] + R.m[3 * i + 2] * Il[5];
        t[3 * i + 2] = R.m[3 * i] * Il[4] + R.m[3 * i + 1] * Il[5] + R.m[3 * i + 2] * Il[2];
    }
    const float cc = dot3(c, c);
    s.I[0] = t[0] * R.m[0] + t[1] * R.m[1] + t[2] * R.m[2] + m * (cc - c.x * c.x);
    s.I[1] = t[3] * R.m[3] + t[4] * R.m[4] + t[5] * R.m[5] + m * (cc - c.y * c.y);
    s.I[2] = t[6] * R.m[6] + t[7] * R.m[7] + t[8] * R.m[8] + m * (cc - c.z * c.z);
    s.I[3] = t[0] * R.m[3] + t[1] * R.m[4] + t[2] * R.m[5] - m * c.x * c.y;
    s.I[4] = t[0] * R.m[6] + t[1] * R.m[7] + t[2] * R.m[8] - m * c.x * c.z;
    s.I[5] = t[3] * R.m[6] + t[4] * R.m[7] + t[5] * R.m[8] - m * c.y * c.z;
    return s;
}

__device__ __forceinline__ m33 quat_to_mat(float w, float x, float y, float z) {
    m33 R;
    R.m[0] = 1 - 2 * (y * y + z * z); R.m[1] = 2 * (x * y - w * z); R.m[2] = 2 * (x * z + w * y);
    R.m[3] = 2 * (x * y + w * z); R.m[4] = 1 - 2 * (x * x + z * z); R.m[5] = 2 * (y * z - w * x);
    R.m[6] = 2 * (x * z - w * y); R.m[7] = 2 * (y * z + w * x); R.m[8] = 1 - 2 * (x * x + y * y);
    return R;
}
__device__ __forceinline__ m33 axis_angle(f3 a, float th) {
    float s, c; fast_sincosf(th, &s, &c);               // |th| stays within the joint limits (< 2 pi): MUFU precision, ~5e-7
    const float t = 1 - c; m33 R;
    R.m[0] = c + a.x * a.x * t;       R.m[1] = a.x * a.y * t - a.z * s; R.m[2] = a.x * a.z * t + a.y * s;
    R.m[3] = a.y * a.x * t + a.z * s; R.m[4] = c + a.y * a.y * t;       R.m[5] = a.y * a.z * t - a.x * s;
    R.m[6] = a.z * a.x * t - a.y * s; R.m[7] = a.z * a.y * t + a.x * s; R.m[8] = c + a.z * a.z * t;
    return R;
}

// terrain height and unit normal under world point (x, y): the triangle of the heightfield cell below it.  `trimesh`
// (terrain.mesh_type "trimesh"): the cell is cut along (i,j)-(i+1,j+1) like the mesh convert_heightfield_to_trimesh builds
// (legged_gym/utils/terrain_utils.py:887-900: triangles (ind0, ind3, ind1) and (ind0, ind2, ind3)); heightfield: along
// (i+1,j)-(i,j+1).  Warp-uniform.
__device__ __forceinline__ void terrain_query(const TerrainDev &tr, float hs, float vs, float border, bool trimesh, float x, float y, float &h, f3 &n) {
    if (tr.hf == nullptr) { h = 0.f; n = mk3(0.f, 0.f, 1.f); return; }
    const float ihs = 1.f / hs;                    // grid coordinates by the reciprocal, like the oracle (same cell on both sides)
    const float gx = (x + border) * ihs, gy = (y + border) * ihs;
    int i = (int)floorf(gx), j = (int)floorf(gy);
    i = max(0, min(i, tr.rows - 2)); j = max(0, min(j, tr.cols - 2));
    const float u = fminf(fmaxf(gx - (float)i, 0.f), 1.f), w = fminf(fmaxf(gy - (float)j, 0.f), 1.f);
    const int16_t *p = tr.hf + (size_t)i * tr.cols + j;
    const float h00 = (float)__ldg(p) * vs, h01 = (float)__ldg(p + 1) * vs;
    const float h10 = (float)__ldg(p + tr.cols) * vs, h11 = (float)__ldg(p + tr.cols + 1) * vs;
    float dhx, dhy;
    if (trimesh) {
        if (u >= w) { dhx = h10 - h00; dhy = h11 - h10; } else { dhx = h11 - h01; dhy = h01 - h00; }
        h = h00 + u * dhx + w * dhy;
    } else if (u + w <= 1.f) { dhx = h10 - h00; dhy = h01 - h00; h = h00 + u * dhx + w * dhy; }
    else { dhx = h11 - h01; dhy = h11 - h10; h = h11 - (1.f - u) * dhx - (1.f - w) * dhy; }
    const f3 g = mk3(-dhx * ihs, -dhy * ihs, 1.f);
    n = g * rsqrtf(dot3(g, g));
}

__device__ __forceinline__ float impedance(const float *tf, float pos) {
    const float x = __fdividef(fabsf(pos), tf[TF_WIDTH]);
    if (x >= 1.f) return tf[TF_DMAX];
    const float mid = tf[TF_MID], p = tf[TF_POWER];
    float y;
    if (p == 2.f) y = x < mid ? __fdividef(x * x, mid) : 1.f - __fdividef((1.f - x) * (1.f - x), 1.f - mid);     // the shipped solimp: no powf
    else if (x < mid) y = powf(x, p) / powf(mid, p - 1.f);
    else y = 1.f - powf(1.f - x, p) / powf(1.f - mid, p - 1.f);
    return tf[TF_D0] + y * (tf[TF_DMAX] - tf[TF_D0]);
}

// symmetric 6x6 stored as lower triangle, index (i>=j): i*(i+1)/2 + j
__device__ __forceinline__ constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// in: S (lower tri, SPD).  out: Sinv (lower tri)
__device__ __forceinline__ void spd6_inverse(const float *S, float *Sinv) {
    float L[21], Li[21], rd[6];        // rd[j] = 1 / L[j][j]
#pragma unroll
    for (int j = 0; j < 6; j++) {
        float d = S[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[tri(j, k)] * L[tri(j, k)];
        const float r = rsqrtf(d);
        rd[j] = r;
        L[tri(j, j)] = d * r;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            float s = S[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[tri(i, k)] * L[tri(j, k)];
            L[tri(i, j)] = s * r;
        }
    }
    // Li = L^-1 (lower)
#pragma unroll
    for (int j = 0; j < 6; j++) {
        Li[tri(j, j)] = rd[j];
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            float s = 0.f;
#pragma unroll
            for (int k = j; k < i; k++) s += L[tri(i, k)] * Li[tri(k, j)];
            Li[tri(i, j)] = -s * rd[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            float s = 0.f;
#pragma unroll
            for (int k = i; k < 6; k++) s += Li[tri(k, i)] * Li[tri(k, j)];
            Sinv[tri(i, j)] = s;
        }
}

// One env, `decimation` substeps.  C chains of 3 joints.  All 32 lanes execute every collective.
template <int C>
__device__ void dynamics_warp(const TaskDev &T, const B200Buffers &B, const TerrainDev &tr, const float *ms, float *ws,
                              const float *actions_in, int env, int lane, int *cost_out, const int mode) {
    const int sim_only = mode & DYN_MODE_SIM_ONLY;
    const bool late_actions = (mode & DYN_MODE_LATE_ACTIONS) != 0;   // `actions_in` is being filled by fetch_actions_kernel (see there)
    const float *tf = T.f;
    const int A = 3 * C, L = T.i[TI_L], NS = T.i[TI_NSPHERES];
    const int *msi = (const int *)(ms + MS_INT);
    const float h = tf[TF_SIM_DT];
    const bool leg = lane < C;
    const int c = leg ? lane : C - 1;       // chain walked by this lane (lanes >= C shadow the last chain, results masked)

#ifdef DYN_TIMING
    const unsigned long long t_begin = dyn_now();
    unsigned dbg_sweeps = 0, dbg_rows = 0;
#endif
    // ---------------- load state ----------------
    float *pd = ws + WS_PD + 16 * c;        // this lane's chain (lanes >= C alias the last chain and write the same values)
    const smaddr_t st_a = sm_addr(ws + WS_ST), qs_a = sm_addr(ws + WS_QS + 12 * c);
    // the reference shifts the base mass / scales the geom friction only when the DR switch is on (set_mass_shift /
    // set_friction_ratio are called from _randomize_* alone, genesis_simulator.py:62-82,665-697); the observation-side
    // buffers start at 1 / 0 (genesis_simulator.py:648) and must not leak into the physics otherwise
    const float env_arm = B.joint_armature[env], env_dmp = B.joint_damping[env], env_fls = B.joint_friction[env];
    if (lane == 0) {
        const float mass_add = T.i[TI_RAND_MASS] ? B.added_mass[env] : 0.f, fric_ratio = T.i[TI_RAND_FRICTION] ? B.friction[env] : 1.f;
        sts128(st_a, B.base_pos[env * 3], B.base_pos[env * 3 + 1], B.base_pos[env * 3 + 2], mass_add);
        sts128(st_a + 16u, B.base_quat_wxyz[env * 4], B.base_quat_wxyz[env * 4 + 1], B.base_quat_wxyz[env * 4 + 2], B.base_quat_wxyz[env * 4 + 3]);
        sts128(st_a + 32u, B.base_lin_w[env * 3], B.base_lin_w[env * 3 + 1], B.base_lin_w[env * 3 + 2], fmaxf(tf[TF_GEOM_MU] * fric_ratio, tf[TF_TERRAIN_MU]));
        sts128(st_a + 48u, B.base_ang_w[env * 3], B.base_ang_w[env * 3 + 1], B.base_ang_w[env * 3 + 2], 0.f);
        sts128(st_a + 64u, B.com_bias[env * 3], B.com_bias[env * 3 + 1], B.com_bias[env * 3 + 2], 0.f);
    }

    for (int e = lane; e < 48; e += 32) ws[WS_WARM + e] = B.contact_warm[env * 48 + e];
    // ---------------- pre-step: last_* (genesis_simulator.py:21-24); the actions are only FETCHED here ----------------
    // Each joint lane starts an asynchronous copy of its action into the (still unused) torque slot of its chain.  Nothing
    // waits for it before the first substep's forward kinematics are done (see "actions arrive" below), so the latency of
    // the read -- a PCIe round trip when b200_env_step hands over pinned HOST memory -- is off every warp's critical path.
    if (lane < A) {
        const int o = env * A + lane;
        if (!late_actions) cp_async_f32(ws + WS_QS + 12 * (lane / 3) + 8 + lane % 3, actions_in + o);
        B.last_dof_vel[o] = B.dof_vel[o];
        if (lane < 3 * T.i[TI_F]) B.last_feet_vel[env * 3 * T.i[TI_F] + lane] = B.feet_vel[env * 3 * T.i[TI_F] + lane];
        if (lane < 3) { B.last_base_lin_vel[env * 3 + lane] = B.base_lin_vel[env * 3 + lane]; B.last_base_ang_vel[env * 3 + lane] = B.base_ang_vel[env * 3 + lane]; }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int j = 3 * c + k, o = env * A + j;
        const float arm_k = T.i[TI_RAND_ARMATURE] ? env_arm : ms[MS_BODY + (1 + j) * B200_BODY_STRIDE + 19];
        const float dmp_k = T.i[TI_RAND_JDAMPING] ? env_dmp : 0.f;
        pd[3 + k] = B.kp_scale[o] * tf[TF_KP]; pd[6 + k] = B.kd_scale[o] * tf[TF_KD];
        pd[9 + k] = arm_k + h * dmp_k; pd[12 + k] = dmp_k;
        if (k == 0) pd[15] = T.i[TI_RAND_JFRICTION] ? env_fls : 0.f;       // frictionloss bound: one value per env
    }
    if (leg) {
        const int o = env * A + 3 * c;
        sts128(qs_a, B.dof_pos[o], B.dof_pos[o + 1], B.dof_pos[o + 2], 0.f);
        sts128(qs_a + 16u, B.dof_vel[o], B.dof_vel[o + 1], B.dof_vel[o + 2], 0.f);
    }
    __syncwarp();

    int cost = 0;                    // solver work of this env in this launch (sweeps x rows, + rows built): feeds the next launches' order
    for (int sub = 0; sub <= T.i[TI_DECIMATION]; sub++) {
        const bool last_pass = sub == T.i[TI_DECIMATION];   // kinematics-only pass for the outputs
        PHASE_SYNC_A();
        // the state this phase needs, from its parking place (live until the smooth accelerations are out)
        float q[3], qd[3], tau[3] = {0.f, 0.f, 0.f}, Qw, Qx, Qy, Qz;
        f3 vb, wb;
        {
            const float4 a4 = lds128(st_a + 16u), b4 = lds128(st_a + 32u), c4 = lds128(st_a + 48u), q4 = lds128(qs_a), v4 = lds128(qs_a + 16u);
            Qw = a4.x; Qx = a4.y; Qy = a4.z; Qz = a4.w; vb = mk3(b4.x, b4.y, b4.z); wb = mk3(c4.x, c4.y, c4.z);
            q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; qd[0] = v4.x; qd[1] = v4.y; qd[2] = v4.z;
        }
        // ---------------- forward kinematics, velocities, RNE, CRBA along this lane's chain ----------------
        const m33 R0 = quat_to_mat(Qw, Qx, Qy, Qz);
        // The chain walk below is serial in the body index and redundant over the lanes (lanes >= C shadow the last chain), so
        // only what MUST follow the chain stays in it: frames, velocities and the velocity-product accelerations.  The body-local
        // part of the RNE -- world inertia about the base origin (R I R^T, parallel axes), momenta, bias wrench: ~180
        // instructions per body -- is done afterwards by ONE LANE PER BODY (lane j = body 1 + j), all bodies at once.
        // Hand-over through the chain's parking slots in a part of the A matrix's storage that nothing else touches before
        // the constraint rows are built: [k][16] per body (first aw, av from the chain lane; then m, h, I, fn, ff from the
        // body lane), then the three moment arms.
        constexpr int PARK = 60;         // 16-byte aligned slots: the body records move as 4 x 16 bytes
        float *park = ws + WS_AM + 640 + c * PARK;       // lanes >= C read the last chain's slot (and write nothing)
        {
            m33 Rp = R0; f3 op = mk3(0.f, 0.f, 0.f), wp = wb, vp = vb;
            f3 awp = mk3(0.f, 0.f, 0.f), avp = cross3(vb, wb) + mk3(0.f, 0.f, tf[TF_GRAV]);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int b = 1 + 3 * c + k;
                const float *Bd = ms + MS_BODY + b * B200_BODY_STRIDE;
                const f3 ax = mk3(Bd[3], Bd[4], Bd[5]);
                const f3 o = op + mul(Rp, mk3(Bd[0], Bd[1], Bd[2]));
                const f3 a = mul(Rp, ax);
                const f3 sv = cross3(o, a);
                const m33 Rk = mul(Rp, axis_angle(ax, q[k]));
                const f3 w = wp + a * qd[k], v = vp + sv * qd[k];
                if (leg) {
                    float *fr = ws + WS_FR + b * 12;
#pragma unroll
                    for (int e = 0; e < 9; e++) fr[e] = Rk.m[e];
                    fr[9] = o.x; fr[10] = o.y; fr[11] = o.z;
                    float *axp = ws + WS_AX + (b - 1) * 3; axp[0] = a.x; axp[1] = a.y; axp[2] = a.z;
                    float *vl = ws + WS_VEL + b * 6; vl[0] = w.x; vl[1] = w.y; vl[2] = w.z; vl[3] = v.x; vl[4] = v.y; vl[5] = v.z;
                    park[48 + 3 * k] = sv.x; park[49 + 3 * k] = sv.y; park[50 + 3 * k] = sv.z;
                }
                if (!last_pass) {
                    const f3 aw = awp + cross3(w, a) * qd[k];
                    const f3 av = avp + (cross3(w, sv) + cross3(v, a)) * qd[k];
                    if (leg) { const smaddr_t pa = sm_addr(park + 16 * k); sts128(pa, aw.x, aw.y, aw.z, av.x); sts128(pa + 16u, av.y, av.z, 0.f, 0.f); }
                    awp = aw; avp = av;
                }
                Rp = Rk; op = o; wp = w; vp = v;
            }
        }
        if (lane == 0) {
            float *fr = ws + WS_FR;
#pragma unroll
            for (int e = 0; e < 9; e++) fr[e] = R0.m[e];
            fr[9] = fr[10] = fr[11] = 0.f;
            float *vl = ws + WS_VEL; vl[0] = wb.x; vl[1] = wb.y; vl[2] = wb.z; vl[3] = vb.x; vl[4] = vb.y; vl[5] = vb.z;
        }
        __syncwarp();
        if (last_pass) break;

        if (sub == 0) {
            // ---------------- actions arrive: clip, action history, delay queue (legged_robot.py:230-252) ----------------
            if (late_actions) grid_dependency_wait();       // the copy kernel launched ahead of this one has finished and is visible
            else cp_async_wait_all();
            __syncwarp();
            float a_applied = 0.f;           // lane j < A: the action joint j is driven with (the clipped action, or a delayed one)
            if (lane < A) {
                const int o = env * A + lane;
                const float raw = late_actions ? __ldcg(actions_in + o) : ws[WS_QS + 12 * (lane / 3) + 8 + lane % 3];
                // sim_only (b200_simulator_step, plugin mode): LeggedRobot._pre_sim_step already clipped / delayed the actions and
                // keeps the action history itself; only GenesisSimulator.step's part runs here
                const float a = sim_only ? raw : fminf(fmaxf(raw, -tf[TF_CLIP_ACTIONS]), tf[TF_CLIP_ACTIONS]);
                if (!sim_only) {
                    B.llast_actions[o] = B.last_actions[o];
                    B.last_actions[o] = B.actions[o];
                    B.actions[o] = a;
                }
                a_applied = a;
                if (T.i[TI_CTRL_DELAY] && !sim_only) {        // legged_robot.py:240-245: push the action into the env's queue, drive with slot action_delay
                    const int depth = T.i[TI_CTRL_DELAY_HI] + 1, dly = min(max(B.action_delay[env], 0), depth - 1);
                    float *qu = B.action_queue + (size_t)env * depth * A + lane;
                    for (int d = depth - 1; d > 0; d--) qu[d * A] = qu[(d - 1) * A];
                    qu[0] = a;
                    a_applied = qu[dly * A];
                }
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float a = __shfl_sync(B200_FULL_MASK, a_applied, 3 * c + k);
                pd[k] = a * tf[TF_ACTION_SCALE] + tf[TF_DEFAULT_DOF_POS + 3 * c + k];
            }
            __syncwarp();                    // every lane has read its raw action before the torques replace them
        }
        // ---------------- PD torque (genesis_simulator.py:630-642) ----------------
#pragma unroll
        for (int k = 0; k < 3; k++) tau[k] = pd[3 + k] * (pd[k] - q[k]) - pd[6 + k] * qd[k];
        if (leg) sts128(qs_a + 32u, tau[0], tau[1], tau[2], 0.f);      // the last substep's is the reported torque

        // ---------------- body-local RNE: one lane per body ----------------
        if (lane < A) {
            const int b = 1 + lane;
            const float *Bd = ms + MS_BODY + b * B200_BODY_STRIDE;
            const smaddr_t fr_a = sm_addr(ws + WS_FR + b * 12), vl_a = sm_addr(ws + WS_VEL + b * 6);
            const smaddr_t pk_a = sm_addr(ws + WS_AM + 640 + (lane / 3) * PARK + 16 * (lane % 3));
            const float4 f0 = lds128(fr_a), f1 = lds128(fr_a + 16u), f2 = lds128(fr_a + 32u);       // R (9), o (3)
            const float2 v0 = lds64(vl_a), v1 = lds64(vl_a + 8u), v2 = lds64(vl_a + 16u);           // w (3), v (3)
            const float4 a0 = lds128(pk_a);                                                         // aw (3), av.x
            const float2 a1 = lds64(pk_a + 16u);                                                    // av.y, av.z
            m33 Rk;
            Rk.m[0] = f0.x; Rk.m[1] = f0.y; Rk.m[2] = f0.z; Rk.m[3] = f0.w; Rk.m[4] = f1.x; Rk.m[5] = f1.y; Rk.m[6] = f1.z; Rk.m[7] = f1.w; Rk.m[8] = f2.x;
            const f3 o = mk3(f2.y, f2.z, f2.w), w = mk3(v0.x, v0.y, v1.x), v = mk3(v1.y, v2.x, v2.y);
            const f3 aw = mk3(a0.x, a0.y, a0.z), av = mk3(a0.w, a1.x, a1.y);
            const f3 cm = o + mul(Rk, mk3(Bd[6], Bd[7], Bd[8]));
            const SIn sik = si_body(Bd[9], cm, Bd + 10, Rk);
            f3 LA, PA, LV, PV;
            si_apply(sik, aw, av, LA, PA);
            si_apply(sik, w, v, LV, PV);
            const f3 fnk = LA + cross3(w, LV) + cross3(v, PV);
            const f3 ffk = PA + cross3(w, PV);
            sts128(pk_a, sik.m, sik.h.x, sik.h.y, sik.h.z);
            sts128(pk_a + 16u, sik.I[0], sik.I[1], sik.I[2], sik.I[3]);
            sts128(pk_a + 32u, sik.I[4], sik.I[5], fnk.x, fnk.y);
            sts128(pk_a + 48u, fnk.z, ffk.x, ffk.y, ffk.z);
        }
        __syncwarp();
        PHASE_SYNC_B();
        // backward pass along the chain: composite inertias, bias forces, CRBA columns
        float Dm[6];           // chain block, lower tri (00,10,11,20,21,22)
        f3 BP[3], BL[3];       // base coupling: column k of B^T = (P_k (lin rows), L_k (ang rows))
        float biasl[3];
        f3 a_[3], sv_[3];       // joint axes (written by this chain's lane before the barrier above) and their moment arms
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float *axp = ws + WS_AX + (3 * c + k) * 3;
            a_[k] = mk3(axp[0], axp[1], axp[2]);
            sv_[k] = mk3(park[48 + 3 * k], park[49 + 3 * k], park[50 + 3 * k]);
        }
        SIn comp;
        f3 fns, ffs;
        const smaddr_t park_a = sm_addr(park);
        {
            const float4 r0 = lds128(park_a + 128u), r1 = lds128(park_a + 144u), r2 = lds128(park_a + 160u), r3 = lds128(park_a + 176u);
            comp.m = r0.x; comp.h = mk3(r0.y, r0.z, r0.w);
            comp.I[0] = r1.x; comp.I[1] = r1.y; comp.I[2] = r1.z; comp.I[3] = r1.w; comp.I[4] = r2.x; comp.I[5] = r2.y;
            fns = mk3(r2.z, r2.w, r3.x); ffs = mk3(r3.y, r3.z, r3.w);
        }
#pragma unroll
        for (int k = 2; k >= 0; k--) {
            if (k < 2) {
                const smaddr_t pa = park_a + 64u * k;
                const float4 r0 = lds128(pa), r1 = lds128(pa + 16u), r2 = lds128(pa + 32u), r3 = lds128(pa + 48u);
                SIn sik;
                sik.m = r0.x; sik.h = mk3(r0.y, r0.z, r0.w);
                sik.I[0] = r1.x; sik.I[1] = r1.y; sik.I[2] = r1.z; sik.I[3] = r1.w; sik.I[4] = r2.x; sik.I[5] = r2.y;
                si_acc(comp, sik); fns = fns + mk3(r2.z, r2.w, r3.x); ffs = ffs + mk3(r3.y, r3.z, r3.w);
            }
            biasl[k] = dot3(a_[k], fns) + dot3(sv_[k], ffs);
            f3 Lk, Pk; si_apply(comp, a_[k], sv_[k], Lk, Pk);
            BP[k] = Pk; BL[k] = Lk;
#pragma unroll
            for (int i2 = 0; i2 <= k; i2++) Dm[tri(k, i2)] = dot3(a_[i2], Lk) + dot3(sv_[i2], Pk);
            Dm[tri(k, k)] += pd[9 + k];
        }
        // chain block inverse (3x3 SPD), G = Dinv B^T
        float Di[6], G[3][6];
        {
            const float i00 = rsqrtf(Dm[0]), l10 = Dm[1] * i00, l20 = Dm[3] * i00;
            const float i11 = rsqrtf(Dm[2] - l10 * l10), l21 = (Dm[4] - l20 * l10) * i11;
            const float i22 = rsqrtf(Dm[5] - l20 * l20 - l21 * l21);
            const float i10 = -l10 * i00 * i11, i21 = -l21 * i11 * i22, i20 = -(l20 * i00 + l21 * i10) * i22;
            Di[0] = i00 * i00 + i10 * i10 + i20 * i20; Di[1] = i10 * i11 + i20 * i21; Di[2] = i11 * i11 + i21 * i21;
            Di[3] = i20 * i22; Di[4] = i21 * i22; Di[5] = i22 * i22;
        }
        float Bt[3][6];
#pragma unroll
        for (int k = 0; k < 3; k++) { Bt[k][0] = BP[k].x; Bt[k][1] = BP[k].y; Bt[k][2] = BP[k].z; Bt[k][3] = BL[k].x; Bt[k][4] = BL[k].y; Bt[k][5] = BL[k].z; }
#pragma unroll
        for (int a2 = 0; a2 < 3; a2++)
#pragma unroll
            for (int e = 0; e < 6; e++) G[a2][e] = Di[tri(a2, 0)] * Bt[0][e] + Di[tri(a2, 1)] * Bt[1][e] + Di[tri(a2, 2)] * Bt[2][e];
        // chain-local part of the smooth acceleration: tl = Dinv (tau_applied - damping qd - bias)
        float tl[3];
        {
            float rl[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float eff = ms[MS_BODY + (1 + 3 * c + k) * B200_BODY_STRIDE + 18];
                rl[k] = fminf(fmaxf(tau[k], -eff), eff) - pd[12 + k] * qd[k] - biasl[k];
            }
#pragma unroll
            for (int a2 = 0; a2 < 3; a2++) tl[a2] = Di[tri(a2, 0)] * rl[0] + Di[tri(a2, 1)] * rl[1] + Di[tri(a2, 2)] * rl[2];
        }
        // D^-1 and G go to their place for the constraint rows now: the base block below does not keep them in registers
        // (it reads G back), which is 24 registers less across the 6x6 inverse
        if (leg) {
            float *mi = ws + WS_MI + c * 24;
#pragma unroll
            for (int e = 0; e < 6; e++) mi[e] = Di[e];
#pragma unroll
            for (int a2 = 0; a2 < 3; a2++)
#pragma unroll
                for (int e = 0; e < 6; e++) mi[6 + a2 * 6 + e] = G[a2][e];
        }
        // Everything the base needs from the chains is a sum over the C chain lanes of 43 per-chain numbers (composite
        // inertia 10, bias force 6, Schur terms B D^-1 B^T 21, B tl 6).  Transposed through shared memory: the chain
        // lanes park their numbers in rows, lane e adds up column e, every lane reads the 43 totals back with vector
        // loads -- ~75 instructions instead of 43 shuffle reductions of 5.  The scratch is the part of the A matrix that
        // does not alias the frames (A is only written after the collision phase).
        constexpr int NRED = 43, RLD = 45, TOT = B200_MAX_CHAINS * RLD;     // totals row: 16-byte aligned for any C
        float *red = ws + WS_AM + 400;
        __syncwarp();
        if (leg) {
            float *row = red + c * RLD;
            row[0] = comp.m; row[1] = comp.h.x; row[2] = comp.h.y; row[3] = comp.h.z;
#pragma unroll
            for (int e = 0; e < 6; e++) row[4 + e] = comp.I[e];
            row[10] = fns.x; row[11] = fns.y; row[12] = fns.z; row[13] = ffs.x; row[14] = ffs.y; row[15] = ffs.z;
#pragma unroll
            for (int i2 = 0; i2 < 6; i2++)
#pragma unroll
                for (int j2 = 0; j2 <= i2; j2++) row[16 + tri(i2, j2)] = Bt[0][i2] * G[0][j2] + Bt[1][i2] * G[1][j2] + Bt[2][i2] * G[2][j2];
#pragma unroll
            for (int e = 0; e < 6; e++) row[37 + e] = Bt[0][e] * tl[0] + Bt[1][e] * tl[1] + Bt[2][e] * tl[2];
        }
        __syncwarp();
        {
            float t0 = 0.f, t1 = 0.f;
#pragma unroll
            for (int cc = 0; cc < C; cc++) { t0 += red[cc * RLD + lane]; if (lane + 32 < NRED) t1 += red[cc * RLD + lane + 32]; }
            red[TOT + lane] = t0;
            if (lane + 32 < NRED) red[TOT + lane + 32] = t1;
        }
        __syncwarp();
        float tot[NRED + 1];
#pragma unroll
        for (int e4 = 0; e4 < (NRED + 1) / 4; e4++) {
            const float4 v = reinterpret_cast<const float4 *>(red + TOT)[e4];
            tot[4 * e4] = v.x; tot[4 * e4 + 1] = v.y; tot[4 * e4 + 2] = v.z; tot[4 * e4 + 3] = v.w;
        }
        // base body + the chain roots
        SIn cb;
        f3 fnb, ffb;
        {
            const float *Bd = ms + MS_BODY;
            const float4 cs4 = lds128(st_a + 64u);                              // per-env COM shift; base mass shift
            const f3 cm = mul(R0, mk3(Bd[6] + cs4.x, Bd[7] + cs4.y, Bd[8] + cs4.z));
            SIn sb = si_body(Bd[9] + lds32(st_a + 12u), cm, Bd + 10, R0);
            const f3 aw = mk3(0.f, 0.f, 0.f), av = cross3(vb, wb) + mk3(0.f, 0.f, tf[TF_GRAV]);
            f3 LA, PA, LV, PV;
            si_apply(sb, aw, av, LA, PA);
            si_apply(sb, wb, vb, LV, PV);
            fnb = LA + cross3(wb, LV) + cross3(vb, PV);
            ffb = PA + cross3(wb, PV);
            cb = sb;
            cb.m += tot[0];
            cb.h.x += tot[1]; cb.h.y += tot[2]; cb.h.z += tot[3];
#pragma unroll
            for (int e = 0; e < 6; e++) cb.I[e] += tot[4 + e];
            fnb.x += tot[10]; fnb.y += tot[11]; fnb.z += tot[12];
            ffb.x += tot[13]; ffb.y += tot[14]; ffb.z += tot[15];
        }
        float S[21];
        {
            // base block: [m 1, [h]x^T ; [h]x, I]  ordering lin(0..2), ang(3..5); minus the Schur terms of the chains
#pragma unroll
            for (int e = 0; e < 21; e++) S[e] = 0.f;
            S[tri(0, 0)] = S[tri(1, 1)] = S[tri(2, 2)] = cb.m;
            S[tri(3, 1)] = -cb.h.z; S[tri(3, 2)] = cb.h.y; S[tri(4, 0)] = cb.h.z; S[tri(4, 2)] = -cb.h.x; S[tri(5, 0)] = -cb.h.y; S[tri(5, 1)] = cb.h.x;
            S[tri(3, 3)] = cb.I[0]; S[tri(4, 4)] = cb.I[1]; S[tri(5, 5)] = cb.I[2]; S[tri(4, 3)] = cb.I[3]; S[tri(5, 3)] = cb.I[4]; S[tri(5, 4)] = cb.I[5];
#pragma unroll
            for (int e = 0; e < 21; e++) S[e] -= tot[16 + e];
        }
        float Sinv[21];
        spd6_inverse(S, Sinv);
        if (lane == 0) {
#pragma unroll
            for (int e = 0; e < 21; e++) ws[WS_MI + 4 * 24 + e] = Sinv[e];
        }
        // smooth acceleration: a = M^-1 (tau_applied - damping qd - bias)
        float ab[6], al[3];
        {
            float rb[6] = {-ffb.x, -ffb.y, -ffb.z, -fnb.x, -fnb.y, -fnb.z};
#pragma unroll
            for (int e = 0; e < 6; e++) rb[e] -= tot[37 + e];
#pragma unroll
            for (int i2 = 0; i2 < 6; i2++) {
                float s = 0.f;
#pragma unroll
                for (int j2 = 0; j2 < 6; j2++) s += Sinv[tri(i2, j2)] * rb[j2];
                ab[i2] = s;
            }
#pragma unroll
            for (int a2 = 0; a2 < 3; a2++) {
                float s = tl[a2];
#pragma unroll
                for (int e = 0; e < 6; e++) s -= ws[WS_MI + c * 24 + 6 + a2 * 6 + e] * ab[e];
                al[a2] = s;
            }
        }
        if (leg) {
#pragma unroll
            for (int k = 0; k < 3; k++) { ws[WS_NU + 6 + 3 * c + k] = qd[k]; ws[WS_AF + 6 + 3 * c + k] = al[k]; }
        }
        if (lane == 0) {
            ws[WS_NU + 0] = vb.x; ws[WS_NU + 1] = vb.y; ws[WS_NU + 2] = vb.z; ws[WS_NU + 3] = wb.x; ws[WS_NU + 4] = wb.y; ws[WS_NU + 5] = wb.z;
#pragma unroll
            for (int e = 0; e < 6; e++) ws[WS_AF + e] = ab[e];
        }

        PHASE_SYNC_C();
        // ---------------- collision detection: two spheres per lane ----------------
        float sdist[2]; f3 sn[2], sx[2]; bool act[2];
        const float4 p4 = lds128(st_a);
        const f3 p = mk3(p4.x, p4.y, p4.z);
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int s = lane + 32 * t;
            act[t] = false; sdist[t] = 0.f; sn[t] = mk3(0.f, 0.f, 1.f); sx[t] = mk3(0.f, 0.f, 0.f);
            if (s < NS) {
                const float *Sp = ms + MS_SPH + 4 * s;
                const float *fr = ws + WS_FR + msi[B200_MAX_LINKS + s] * 12;
                m33 Rb;
#pragma unroll
                for (int e = 0; e < 9; e++) Rb.m[e] = fr[e];
                sx[t] = mk3(fr[9], fr[10], fr[11]) + mul(Rb, mk3(Sp[0], Sp[1], Sp[2]));
                float hh;
                terrain_query(tr, tf[TF_HSCALE], tf[TF_VSCALE], tf[TF_BORDER], T.i[TI_TRIMESH] != 0, p.x + sx[t].x, p.y + sx[t].y, hh, sn[t]);
                sdist[t] = (p.z + sx[t].z - hh) * sn[t].z - Sp[3];
                act[t] = sdist[t] < 0.f;
            }
        }
        unsigned m0 = __ballot_sync(B200_FULL_MASK, act[0]), m1 = __ballot_sync(B200_FULL_MASK, act[1]);
        int nact = __popc(m0) + __popc(m1);
        while (nact > B200_KMAX) {   // drop the shallowest contact (ties: highest sphere index)
            float best = act[1] ? sdist[1] : (act[0] ? sdist[0] : -1e30f);
            int bidx = act[1] ? lane + 32 : (act[0] ? lane : -1);
            if (act[0] && act[1] && sdist[0] > sdist[1]) { best = sdist[0]; bidx = lane; }
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) {
                const float ob = __shfl_xor_sync(B200_FULL_MASK, best, o2);
                const int oi = __shfl_xor_sync(B200_FULL_MASK, bidx, o2);
                if (ob > best || (ob == best && oi > bidx)) { best = ob; bidx = oi; }
            }
            if (bidx == lane) act[0] = false;
            if (bidx == lane + 32) act[1] = false;
            m0 = __ballot_sync(B200_FULL_MASK, act[0]); m1 = __ballot_sync(B200_FULL_MASK, act[1]);
            nact--;
        }
        const int nc = nact;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int t = 0; t < 2; t++) {
            if (act[t]) {
                const int s = lane + 32 * t;
                const int slot = t == 0 ? __popc(m0 & lt) : __popc(m0) + __popc(m1 & lt);
                float *ct = ws + WS_CT + slot * 12;
                const float rad = ms[MS_SPH + 4 * s + 3];
                const f3 xc = sx[t] - sn[t] * (rad + 0.5f * sdist[t]);
                ct[0] = xc.x; ct[1] = xc.y; ct[2] = xc.z; ct[3] = sn[t].x; ct[4] = sn[t].y; ct[5] = sn[t].z; ct[6] = sdist[t];
                ct[7] = __int_as_float(msi[B200_MAX_LINKS + s]);                        // body
                ct[8] = __int_as_float(msi[B200_MAX_LINKS + B200_MAX_SPHERES + s]);     // reporting link
                ct[9] = (float)(s + 1);                                                 // warm-start key
            }
        }
        // ---------------- aux rows: joint limits first, then frictionloss, joint order ----------------
        unsigned limmask = 0, flsmask = 0;
        float limpos[3], limsign[3];
        const float4 qa4 = lds128(qs_a);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float *Bd = ms + MS_BODY + (1 + 3 * c + k) * B200_BODY_STRIDE;
            const float qk = k == 0 ? qa4.x : (k == 1 ? qa4.y : qa4.z);
            const bool lo_v = qk < Bd[16], hi_v = qk > Bd[17];
            limsign[k] = lo_v ? 1.f : -1.f;
            limpos[k] = lo_v ? qk - Bd[16] : Bd[17] - qk;
            const unsigned bl = __ballot_sync(B200_FULL_MASK, leg && (lo_v || hi_v));
            const unsigned bf = __ballot_sync(B200_FULL_MASK, leg && pd[15] > 0.f);
#pragma unroll
            for (int cc = 0; cc < C; cc++) { limmask |= ((bl >> cc) & 1u) << (3 * cc + k); flsmask |= ((bf >> cc) & 1u) << (3 * cc + k); }
        }
        const int nlim = min(__popc(limmask), B200_AUXMAX);
        const int nfls = min(__popc(flsmask), B200_AUXMAX - nlim);
        if (leg) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int j = 3 * c + k;
                const unsigned below = (1u << j) - 1u;
                if ((limmask >> j) & 1u) {
                    const int slot = __popc(limmask & below);
                    if (slot < nlim) { float *ax = ws + WS_AUX + 4 * slot; ax[0] = __int_as_float(j); ax[1] = limsign[k]; ax[2] = limpos[k]; ax[3] = -1.f; }
                }
                if ((flsmask >> j) & 1u) {
                    const int slot = nlim + __popc(flsmask & below);
                    if (slot < nlim + nfls) { float *ax = ws + WS_AUX + 4 * slot; ax[0] = __int_as_float(j); ax[1] = 1.f; ax[2] = 0.f; ax[3] = pd[15]; }
                }
            }
        }
        __syncwarp();
        const int R = 3 * nc + nlim + nfls;

        PHASE_SYNC_D();
        // ---------------- constraint rows: one per lane ----------------
        float Jb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, Jl[3] = {0.f, 0.f, 0.f};
        int cl = -1, kind = 5;          // 0 normal, 1/2 tangents, 3 limit, 4 frictionloss, 5 unused
        float rpos = 0.f, bound = 0.f;
        f3 dir = mk3(0.f, 0.f, 0.f);
        int rlink = 0;
        float wkey = 0.f;               // warm-start key of this row (contact: sphere id + 1, aux: code + 1)
        if (lane < 3 * nc) {
            const float *ct = ws + WS_CT + (lane / 3) * 12;
            const int d = lane - 3 * (lane / 3);
            const f3 xc = mk3(ct[0], ct[1], ct[2]), n = mk3(ct[3], ct[4], ct[5]);
            const f3 e = fabsf(n.x) < 0.9f ? mk3(1.f, 0.f, 0.f) : mk3(0.f, 1.f, 0.f);
            f3 t1 = e - n * dot3(e, n); t1 = t1 * rsqrtf(dot3(t1, t1));
            const f3 t2 = cross3(n, t1);
            dir = d == 0 ? n : (d == 1 ? t1 : t2);
            kind = d; rpos = d == 0 ? ct[6] : 0.f; wkey = ct[9];
            const int b = __float_as_int(ct[7]); rlink = __float_as_int(ct[8]);
            Jb[0] = dir.x; Jb[1] = dir.y; Jb[2] = dir.z;
            const f3 xd = cross3(xc, dir); Jb[3] = xd.x; Jb[4] = xd.y; Jb[5] = xd.z;
            if (b > 0) {
                cl = (b - 1) / 3;
                const int kd2 = (b - 1) - 3 * cl;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (k <= kd2) {
                        const int bj = 1 + 3 * cl + k;
                        const float *fr = ws + WS_FR + bj * 12; const float *axp = ws + WS_AX + (bj - 1) * 3;
                        Jl[k] = dot3(dir, cross3(mk3(axp[0], axp[1], axp[2]), xc - mk3(fr[9], fr[10], fr[11])));
                    }
                }
            }
        } else if (lane < R) {
            const float *ax = ws + WS_AUX + 4 * (lane - 3 * nc);
            const int j = __float_as_int(ax[0]);
            cl = j / 3;
            const int k2 = j - 3 * cl;
#pragma unroll
            for (int k = 0; k < 3; k++) Jl[k] = (k == k2) ? ax[1] : 0.f;
            rpos = ax[2]; bound = ax[3];
            kind = ax[3] < 0.f ? 3 : 4;
            wkey = (float)(8 * j + (kind == 3 ? (ax[1] > 0.f ? 7 : 6) : 4) + 1);
        }
        // y = M^-1 J^T for this lane's row, in the arrowhead factorisation M = [D B; B^T M_bb]:
        //   rb = Jb - G_c^T Jl   (c the row's chain, G = D^-1 B^T)      yb = S^-1 rb      yl_c' = [c' == c] D_c^-1 Jl - G_c' yb
        // Only yb and ul = D_c^-1 Jl are formed.  The chain parts yl_c' of the other chains are never needed explicitly:
        //   A[r][s] = J_r . y_s = rb_r . yb_s + [c_r == c_s] Jl_r . ul_s          (the rows stored for the others hold rb, not Jb)
        //   sum_r yl_r,c' f_r = D_c'^-1 (sum_{r on c'} Jl_r f_r) - G_c' (sum_r yb_r f_r)      (after the solve, once per chain)
        // which takes 4 x 18 multiply-adds, their loads of G and twelve registers that lived through the whole solve off
        // every row.
        float Yb[6], ul[3] = {0.f, 0.f, 0.f};
        float vel = 0.f, ja = 0.f;
        {
            float rb[6];
#pragma unroll
            for (int e = 0; e < 6; e++) rb[e] = Jb[e];
            if (cl >= 0) {
                float mi[24];                    // D^-1 (6) and G (18) of the row's chain: six 16-byte loads
                const smaddr_t mi_a = sm_addr(ws + WS_MI + cl * 24);
#pragma unroll
                for (int e = 0; e < 6; e++) { const float4 v4 = lds128(mi_a + 16u * e); mi[4 * e] = v4.x; mi[4 * e + 1] = v4.y; mi[4 * e + 2] = v4.z; mi[4 * e + 3] = v4.w; }
#pragma unroll
                for (int a2 = 0; a2 < 3; a2++) ul[a2] = mi[tri(a2, 0)] * Jl[0] + mi[tri(a2, 1)] * Jl[1] + mi[tri(a2, 2)] * Jl[2];
#pragma unroll
                for (int e = 0; e < 6; e++) rb[e] -= mi[6 + e] * Jl[0] + mi[12 + e] * Jl[1] + mi[18 + e] * Jl[2];
            }
            float Sv[24];                        // the Schur inverse back from shared memory (21 registers not carried through
            {                                    // the collision / aux-row phases): six 16-byte broadcast loads
                const smaddr_t sv_a = sm_addr(ws + WS_MI + 4 * 24);
#pragma unroll
                for (int e = 0; e < 6; e++) { const float4 v4 = lds128(sv_a + 16u * e); Sv[4 * e] = v4.x; Sv[4 * e + 1] = v4.y; Sv[4 * e + 2] = v4.z; Sv[4 * e + 3] = v4.w; }
            }
#pragma unroll
            for (int i2 = 0; i2 < 6; i2++) {
                float s = 0.f;
#pragma unroll
                for (int j2 = 0; j2 < 6; j2++) s += Sv[tri(i2, j2)] * rb[j2];
                Yb[i2] = s;
            }
            float *jr = ws + WS_JR + lane * 10;
#pragma unroll
            for (int e = 0; e < 6; e++) jr[e] = rb[e];
            jr[6] = Jl[0]; jr[7] = Jl[1]; jr[8] = Jl[2]; jr[9] = __int_as_float(cl);
#pragma unroll
            for (int e = 0; e < 6; e++) { vel += Jb[e] * ws[WS_NU + e]; ja += Jb[e] * ws[WS_AF + e]; }
            if (cl >= 0) {
#pragma unroll
                for (int k = 0; k < 3; k++) { vel += Jl[k] * ws[WS_NU + 6 + 3 * cl + k]; ja += Jl[k] * ws[WS_AF + 6 + 3 * cl + k]; }
            }
        }
        __syncwarp();
        // this lane's column of A = J M^-1 J^T
        float Arr = 1.f;
        // contact rows: the three rows of a sphere sit on the same chain, whose Y columns are picked once per contact
        for (int c2 = 0; c2 < nc; c2++) {
            const float *jc = ws + WS_JR + 30 * c2;
            const bool same = __float_as_int(jc[9]) == cl;          // rows on this lane's chain also meet through D_c^-1
            const float y0 = same ? ul[0] : 0.f, y1 = same ? ul[1] : 0.f, y2 = same ? ul[2] : 0.f;
#pragma unroll
            for (int d = 0; d < 3; d++) {
                // one row of J, the same for every lane: five 8-byte broadcast loads (rows are 40 bytes apart, WS_JR is 16-byte aligned)
                const float2 *jr2 = reinterpret_cast<const float2 *>(jc + 10 * d);
                const float2 j01 = jr2[0], j23 = jr2[1], j45 = jr2[2], j67 = jr2[3], j89 = jr2[4];
                float a2 = 0.f;
                a2 += j01.x * Yb[0]; a2 += j01.y * Yb[1]; a2 += j23.x * Yb[2]; a2 += j23.y * Yb[3]; a2 += j45.x * Yb[4]; a2 += j45.y * Yb[5];
                a2 += j67.x * y0 + j67.y * y1 + j89.x * y2;
                ws[WS_AM + (3 * c2 + d) * 33 + lane] = a2;
                if (3 * c2 + d == lane) Arr = a2;
            }
        }
        for (int r = 3 * nc; r < R; r++) {     // auxiliary rows (joint limits, friction loss): any chain
            const float2 *jr2 = reinterpret_cast<const float2 *>(ws + WS_JR + r * 10);
            const float2 j01 = jr2[0], j23 = jr2[1], j45 = jr2[2], j67 = jr2[3], j89 = jr2[4];
            float a2 = 0.f;
            a2 += j01.x * Yb[0]; a2 += j01.y * Yb[1]; a2 += j23.x * Yb[2]; a2 += j23.y * Yb[3]; a2 += j45.x * Yb[4]; a2 += j45.y * Yb[5];
            if (__float_as_int(j89.y) == cl) a2 += j67.x * ul[0] + j67.y * ul[1] + j89.x * ul[2];
            ws[WS_AM + r * 33 + lane] = a2;
            if (r == lane) Arr = a2;
        }
        float imp = (kind == 0 || kind == 3 || kind == 4) ? impedance(tf, rpos) : 0.f;
        {
            const int src = (kind == 1) ? lane - 1 : ((kind == 2) ? lane - 2 : lane);
            imp = __shfl_sync(B200_FULL_MASK, imp, src);
        }
        float f = 0.f, Rr = 0.f, idd = 1.f, wres = 0.f;
        if (lane < 3 * nc) {                  // warm start: same sphere as in the previous substep keeps its force
            const int d = lane - 3 * (lane / 3);
#pragma unroll
            for (int k = 0; k < B200_KMAX; k++) if (ws[WS_WARM + 4 * k] == wkey) f = ws[WS_WARM + 4 * k + 1 + d];
        } else if (lane < R) {
#pragma unroll
            for (int k = 0; k < B200_AUXMAX; k++) if (ws[WS_WARM + 32 + 2 * k] == wkey) f = ws[WS_WARM + 32 + 2 * k + 1];
        }
        ws[WS_FV + lane * 4] = lane < R ? f : 0.f;
        __syncwarp();
        if (lane < R) {
            const float kk = 1.f / (tf[TF_DMAX] * tf[TF_DMAX] * tf[TF_TC] * tf[TF_TC] * tf[TF_DAMPRATIO] * tf[TF_DAMPRATIO]);
            const float bd = 2.f / (tf[TF_DMAX] * tf[TF_TC]);
            const float aref = -bd * vel - kk * imp * rpos;
            Rr = __fdividef(1.f - imp, imp) * Arr;
            idd = __fdividef(1.f, Arr + Rr);
            wres = ja - aref;
            for (int s2 = 0; s2 < R; s2++) wres += ws[WS_AM + s2 * 33 + lane] * ws[WS_FV + s2 * 4];
        }
        __syncwarp();
        // ---------------- projected block Gauss-Seidel with friction-cone projection ----------------
        // One row per lane.  A sweep visits the contacts in order and solves each contact's three rows (normal, two
        // tangents) TOGETHER: with g = w + R f the full residual of a row (w = J a - aref + A f kept incrementally per
        // lane, R the regulariser), the contact's forces move by -(A_cc + R_c)^-1 g_c, then f_n >= 0 and the tangential
        // pair is projected onto the friction disc |f_t| <= mu f_n.  The 3x3 block inverse is formed once per substep;
        // in the sweep every lane evaluates the whole block update redundantly from three shuffled residuals and broadcast
        // shared-memory reads (block inverse, the contact's current forces), so the three rows cost ONE shuffle latency
        // instead of three, the projection needs no further exchange, and every lane folds the three force changes
        // into its own residual with its entries of the (symmetric) rows of A.  Against the row-by-row sweep this takes
        // 0.63x the sweeps to the same tolerance (9.5 instead of 15.1 per substep on the go2_ts steady state) on a
        // dependency chain ~0.65x as long per contact.  Joint-limit / frictionloss rows stay row-wise.
        const int iters = T.i[TI_PGS_ITERS];
        const float *Arow = ws + WS_AM + lane;            // A[r][lane] at Arow[33 * r]
        const bool clamp0 = (kind == 0 || kind == 3);
        const float fhi = kind == 4 ? bound : 3.0e38f, flo = clamp0 ? 0.f : -fhi;
        const float c1n = -(Rr * idd), iddn = -idd;
        const float mu = lds32(st_a + 44u);               // max(terrain mu, geom mu x friction ratio) of this env
        float *blk = ws + WS_JR;                          // per contact 16 floats: b00 b01 b02 b11 | b12 b22 - - | f0 f1 f2 - | f0 f1 f2 -
                                                          // (the forces twice: sweep k reads copy k & 1 and writes the other, so no lane can
                                                          // read a force its first lane has already replaced)
        {   // block inverses: lane 3c + d computes (redundantly with its two neighbours) the inverse of contact c's block
            const int r0 = min(lane - lane % 3, 29);
            const float R0 = __shfl_sync(B200_FULL_MASK, Rr, r0), R1 = __shfl_sync(B200_FULL_MASK, Rr, r0 + 1), R2 = __shfl_sync(B200_FULL_MASK, Rr, r0 + 2);
            __syncwarp();                                  // every lane is done with its reads of the J rows (A is complete)
            if (lane < 3 * nc) {
                const float *Ab = ws + WS_AM + r0 * 33 + r0;
                const float a00 = Ab[0] + R0, a01 = Ab[1], a02 = Ab[2], a11 = Ab[34] + R1, a12 = Ab[35], a22 = Ab[68] + R2;
                const float c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
                const float idet = __fdividef(1.f, a00 * c00 + a01 * c01 + a02 * c02);
                if (lane == r0) {
                    float *q = blk + 16 * (r0 / 3);
                    q[0] = c00 * idet; q[1] = c01 * idet; q[2] = c02 * idet; q[3] = (a00 * a22 - a02 * a02) * idet;
                    q[4] = (a01 * a02 - a00 * a12) * idet; q[5] = (a00 * a11 - a01 * a01) * idet;
                }
                blk[16 * (r0 / 3) + 8 + (lane - r0)] = f;  // the contact's (warm-started) forces, copy 0
            }
            __syncwarp();
        }
        // one address register for the contact blocks, one for this lane's column of A; everything else is immediates
        const smaddr_t blk_a = sm_addr(blk), arow_a = sm_addr(Arow);
        const smaddr_t own_a = blk_a + 64u * (unsigned)(lane / 3) + 4u * (unsigned)(lane % 3);   // this lane's force inside its contact block
#if DYN_PGS_AREG > 0
        // this lane's entries of A for the first DYN_PGS_AREG contacts live in registers over the sweeps (typical contact counts are
        // covered; the others are re-read from shared memory in every sweep): 4 -> step 0.1700 -> 0.1680 ms, 6 -> no gain (registers)
        float areg[DYN_PGS_AREG][3];
#pragma unroll
        for (int c2 = 0; c2 < DYN_PGS_AREG; c2++)
#pragma unroll
            for (int d = 0; d < 3; d++) areg[c2][d] = c2 < nc ? lds32(arow_a + 396u * c2 + 132u * d) : 0.f;
#endif
        int it = 0;
        for (; it < iters; it++) {
            const float fprev = f;
            const unsigned rd = 32u + 16u * (unsigned)(it & 1), wr = 48u - 16u * (unsigned)(it & 1);   // force copy read / written by this sweep
#ifndef DYN_PGS_ROLLED
            // unrolled over the (at most B200_KMAX) contacts: the shuffle sources and every shared-memory offset are
            // immediates, no address / lane arithmetic and one predicate per contact instead of a counted loop
            const smaddr_t bar_ = blk_a + rd, baw_ = blk_a + wr;
#pragma unroll
            for (int c2 = 0; c2 < B200_KMAX; c2++) {
                if (c2 >= nc) break;
                const float g = fmaf(Rr, f, wres);
                const float g0 = __shfl_sync(B200_FULL_MASK, g, 3 * c2), g1 = __shfl_sync(B200_FULL_MASK, g, 3 * c2 + 1), g2 = __shfl_sync(B200_FULL_MASK, g, 3 * c2 + 2);
                const float4 q0 = lds128(blk_a + 64u * c2), fc = lds128(bar_ + 64u * c2);
                const float2 q1 = lds64(blk_a + 64u * c2 + 16u);
                // the first shuffle to arrive feeds the innermost product of each chain
                float n0 = fmaf(-q0.z, g2, fmaf(-q0.y, g1, fmaf(-q0.x, g0, fc.x)));
                float n1 = fmaf(-q1.x, g2, fmaf(-q0.w, g1, fmaf(-q0.y, g0, fc.y)));
                float n2 = fmaf(-q1.y, g2, fmaf(-q1.x, g1, fmaf(-q0.z, g0, fc.z)));
                n0 = fmaxf(n0, 0.f);
                const float lim = mu * n0, t2 = fmaf(n1, n1, n2 * n2);
                if (t2 > lim * lim) { const float sc = lim * fast_rsqrtf(fmaxf(t2, 1e-30f)); n1 *= sc; n2 *= sc; }      // warp-uniform
#if DYN_PGS_AREG > 0
                if (c2 < DYN_PGS_AREG) {
                    wres = fmaf(areg[c2][0], n0 - fc.x, wres);
                    wres = fmaf(areg[c2][1], n1 - fc.y, wres);
                    wres = fmaf(areg[c2][2], n2 - fc.z, wres);
                } else
#endif
                {
                    wres = fmaf(lds32(arow_a + 396u * c2), n0 - fc.x, wres);
                    wres = fmaf(lds32(arow_a + 396u * c2 + 132u), n1 - fc.y, wres);
                    wres = fmaf(lds32(arow_a + 396u * c2 + 264u), n2 - fc.z, wres);
                }
                if (lane == 3 * c2) sts128(baw_ + 64u * c2, n0, n1, n2, 0.f);
            }
#else
            smaddr_t ba = blk_a, aa = arow_a;
            int src = 0;
            for (int c2 = 0; c2 < nc; c2++, ba += 64u, aa += 396u, src += 3) {
                const float g = fmaf(Rr, f, wres);
                const float g0 = __shfl_sync(B200_FULL_MASK, g, src), g1 = __shfl_sync(B200_FULL_MASK, g, src + 1), g2 = __shfl_sync(B200_FULL_MASK, g, src + 2);
                const float4 q0 = lds128(ba), fc = lds128(ba + rd);
                const float2 q1 = lds64(ba + 16u);
                float n0 = fmaf(-q0.x, g0, fmaf(-q0.y, g1, fmaf(-q0.z, g2, fc.x)));
                float n1 = fmaf(-q0.y, g0, fmaf(-q0.w, g1, fmaf(-q1.x, g2, fc.y)));
                float n2 = fmaf(-q0.z, g0, fmaf(-q1.x, g1, fmaf(-q1.y, g2, fc.z)));
                n0 = fmaxf(n0, 0.f);
                const float lim = mu * n0, t2 = fmaf(n1, n1, n2 * n2);
                if (t2 > lim * lim) { const float sc = lim * fast_rsqrtf(fmaxf(t2, 1e-30f)); n1 *= sc; n2 *= sc; }      // warp-uniform
                wres = fmaf(lds32(aa), n0 - fc.x, wres);
                wres = fmaf(lds32(aa + 132u), n1 - fc.y, wres);
                wres = fmaf(lds32(aa + 264u), n2 - fc.z, wres);
                if (lane == src) sts128(ba + wr, n0, n1, n2, 0.f);
            }
#endif
            __syncwarp();                                  // the forces written by the contacts' first lanes ...
            if (lane < 3 * nc) f = lds32(own_a + wr);      // ... are this sweep's result for the lanes that own those rows
            // joint-limit (f >= 0) and frictionloss (|f| <= bound) rows
            const float cf = c1n * f, dlo = flo - f, dhi = fhi - f;     // a lane's force changes only when its own row is visited
            for (int r = 3 * nc; r < R; r++) {
                const float dl = fminf(fmaxf(fmaf(iddn, wres, cf), dlo), dhi);
                const float delta = __shfl_sync(B200_FULL_MASK, dl, r);
                if (lane == r) f += dl;
                wres = fmaf(Arow[33 * r], delta, wres);
            }
            // convergence: largest change of any row over the sweep relative to the largest force (warp-uniform exit)
            const float dmax = warp_max_nonneg(fabsf(f - fprev)), fmx = warp_max_nonneg(fabsf(f));
#ifdef DYN_TIMING
            dbg_sweeps++; dbg_rows += R;
#endif
            if (dmax <= tf[TF_PGS_TOL] * (1.f + fmx)) break;
        }
        cost += (min(it + 1, iters) + 4) * R;          // sweeps made (+ the work of building the rows)
        __syncwarp();
        for (int e = lane; e < 48; e += 32) ws[WS_WARM + e] = 0.f;
        __syncwarp();
        if (lane < 3 * nc) {
            const int c2 = lane / 3, d = lane - 3 * c2;
            if (d == 0) ws[WS_WARM + 4 * c2] = wkey;
            ws[WS_WARM + 4 * c2 + 1 + d] = f;
        } else if (lane < R) {
            const int a2 = lane - 3 * nc;
            ws[WS_WARM + 32 + 2 * a2] = wkey; ws[WS_WARM + 32 + 2 * a2 + 1] = f;
        }
        PHASE_SYNC_E();
        // ---------------- total acceleration, contact forces, integration ----------------
        // a += M^-1 J^T f.  Base part: sum_r yb_r f_r.  Chain part: D_c^-1 tau_c - G_c (sum_r yb_r f_r) with the joint-space force
        // tau_c = sum over the rows on chain c of Jl_r f_r (see the row assembly).  Every lane parks its 6 + 3C products in the
        // (now dead) A matrix, lane e adds up column e over the R rows, the sums go back through shared memory -- 6 + 3C
        // column sums for ~2R instructions instead of 6 + 3C butterfly reductions of 10
        float accb[6], accl[3];
        {
            constexpr int NV = 6 + 3 * C, LD = NV + 1;        // odd row stride: conflict-free column walks
            float *yf = ws + WS_AM;
            __syncwarp();                                      // every lane is done with its reads of A
            if (lane < R) {
#pragma unroll
                for (int e = 0; e < 6; e++) yf[lane * LD + e] = Yb[e] * f;
#pragma unroll
                for (int l2 = 0; l2 < C; l2++)
#pragma unroll
                    for (int k = 0; k < 3; k++) yf[lane * LD + 6 + 3 * l2 + k] = (l2 == cl) ? Jl[k] * f : 0.f;
            }
            __syncwarp();
            float colsum = 0.f;
            if (lane < NV) for (int r = 0; r < R; r++) colsum += yf[r * LD + lane];
            __syncwarp();
            if (lane < NV) yf[lane] = colsum;                  // row 0: sum_r yb_r f_r (6), then tau_c (3 per chain)
            __syncwarp();
            float sb[6];
#pragma unroll
            for (int e = 0; e < 6; e++) { sb[e] = yf[e]; accb[e] = sb[e] + ws[WS_AF + e]; }
            const float t0 = yf[6 + 3 * c], t1 = yf[7 + 3 * c], t2 = yf[8 + 3 * c];
            float mi[24];
            const smaddr_t mi_a = sm_addr(ws + WS_MI + c * 24);
#pragma unroll
            for (int e = 0; e < 6; e++) { const float4 v4 = lds128(mi_a + 16u * e); mi[4 * e] = v4.x; mi[4 * e + 1] = v4.y; mi[4 * e + 2] = v4.z; mi[4 * e + 3] = v4.w; }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float a = ws[WS_AF + 6 + 3 * c + k] + mi[tri(k, 0)] * t0 + mi[tri(k, 1)] * t1 + mi[tri(k, 2)] * t2;
#pragma unroll
                for (int e = 0; e < 6; e++) a -= mi[6 + k * 6 + e] * sb[e];
                accl[k] = a;
            }
        }
        // net contact force per link: what get_links_net_contact_force returns after the decimation loop
        // (genesis_simulator.py:51) is the LAST substep's, so only that one is assembled
        if (sub == T.i[TI_DECIMATION] - 1) {
            {
                float *fv = ws + WS_FV + lane * 4;
                const float fz = lane < 3 * nc ? f : 0.f;
                fv[0] = fz * dir.x; fv[1] = fz * dir.y; fv[2] = fz * dir.z; fv[3] = __int_as_float(lane < 3 * nc ? rlink : -1);
            }
            __syncwarp();
            for (int e = lane; e < 3 * L; e += 32) ws[WS_LF + e] = 0.f;
            __syncwarp();
            if (lane < 3) {                    // contact order, one component per lane: deterministic accumulation per link
                for (int c2 = 0; c2 < nc; c2++) {
                    const float *fv = ws + WS_FV + 12 * c2;
                    ws[WS_LF + 3 * __float_as_int(fv[3]) + lane] += fv[lane] + fv[4 + lane] + fv[8 + lane];
                }
            }
        }
        // semi-implicit Euler on the parked state: reload, advance, park again
        {
            const float4 p4i = lds128(st_a), a4 = lds128(st_a + 16u), b4 = lds128(st_a + 32u), c4 = lds128(st_a + 48u), q4 = lds128(qs_a), v4 = lds128(qs_a + 16u);
            const f3 vn = mk3(b4.x, b4.y, b4.z) + mk3(accb[0], accb[1], accb[2]) * h;
            const f3 wn3 = mk3(c4.x, c4.y, c4.z) + mk3(accb[3], accb[4], accb[5]) * h;
            const f3 pn = mk3(p4i.x, p4i.y, p4i.z) + vn * h;
            const float qd0 = v4.x + h * accl[0], qd1 = v4.y + h * accl[1], qd2 = v4.z + h * accl[2];
            const float wn = sqrtf(dot3(wn3, wn3));
            float dw = 1.f, dx = 0.f, dy = 0.f, dz = 0.f;
            if (wn > 1e-12f) {
                float sn2, cs2; fast_sincosf(0.5f * wn * h, &sn2, &cs2);
                const float s2 = __fdividef(sn2, wn); dw = cs2; dx = wn3.x * s2; dy = wn3.y * s2; dz = wn3.z * s2;
            }
            const float nw = dw * a4.x - dx * a4.y - dy * a4.z - dz * a4.w;
            const float nx = dw * a4.y + dx * a4.x + dy * a4.w - dz * a4.z;
            const float ny = dw * a4.z - dx * a4.w + dy * a4.x + dz * a4.y;
            const float nz = dw * a4.w + dx * a4.z - dy * a4.y + dz * a4.x;
            const float inv = rsqrtf(nw * nw + nx * nx + ny * ny + nz * nz);
            __syncwarp();                                  // every lane has read the old state
            if (lane == 0) {
                sts128(st_a, pn.x, pn.y, pn.z, p4i.w);
                sts128(st_a + 16u, nw * inv, nx * inv, ny * inv, nz * inv);
                sts128(st_a + 32u, vn.x, vn.y, vn.z, b4.w);
                sts128(st_a + 48u, wn3.x, wn3.y, wn3.z, 0.f);
            }
            if (leg) {
                sts128(qs_a, q4.x + h * qd0, q4.y + h * qd1, q4.z + h * qd2, 0.f);
                sts128(qs_a + 16u, qd0, qd1, qd2, 0.f);
            }
        }
        __syncwarp();
    }

#ifdef DYN_TIMING
    if (lane == 0 && env < 65536) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_dyn_timing[env * 4] = t_begin; g_dyn_timing[env * 4 + 1] = dyn_now();
        g_dyn_timing[env * 4 + 2] = dbg_sweeps | ((unsigned long long)smid << 32); g_dyn_timing[env * 4 + 3] = dbg_rows | ((unsigned long long)blockIdx.x << 32);
    }
#endif
    if (cost_out && lane == 0) *cost_out = cost;
    // ---------------- write back (state, frames and velocities of the final state are in shared memory) ----------------
    const float4 pf4 = lds128(st_a), qf4 = lds128(qs_a), vf4 = lds128(qs_a + 16u), tf4 = lds128(qs_a + 32u);
    const f3 p = mk3(pf4.x, pf4.y, pf4.z);
    const float qd[3] = {vf4.x, vf4.y, vf4.z};
    // failure containment (no reference counterpart): a state that left the finite range is NOT written
    // back -- the env keeps the pose it entered the step with, at rest, and is flagged for the env kernel to reset
    bool finite_l = true;
    if (lane < 13) finite_l = isfinite(ws[WS_ST + lane + (lane >= 3) + (lane >= 10)]);   // p | Q | vb | wb (the 4th words are parameters)
    if (leg) finite_l = finite_l && isfinite(qf4.x + qf4.y + qf4.z + vf4.x + vf4.y + vf4.z);
    if (__ballot_sync(B200_FULL_MASK, !finite_l) != 0u) {
        if (lane < 3) { B.base_lin_w[env * 3 + lane] = 0.f; B.base_ang_w[env * 3 + lane] = 0.f; }
        if (lane < A) { B.dof_vel[env * A + lane] = 0.f; B.torques[env * A + lane] = 0.f; B.last_dof_vel[env * A + lane] = 0.f; }
        if (lane < 3 * T.i[TI_F]) { B.feet_vel[env * 3 * T.i[TI_F] + lane] = 0.f; B.last_feet_vel[env * 3 * T.i[TI_F] + lane] = 0.f; }
        for (int e = lane; e < 3 * L; e += 32) B.link_contact_forces[env * 3 * L + e] = 0.f;
        for (int e = lane; e < 48; e += 32) B.contact_warm[env * 48 + e] = 0.f;
        if (lane == 0 && B.nonfinite) B.nonfinite[env] = 1;
        return;
    }
    if (lane < 3) {
        B.base_pos[env * 3 + lane] = ws[WS_ST + lane];
        B.base_lin_w[env * 3 + lane] = ws[WS_ST + 8 + lane];
        B.base_ang_w[env * 3 + lane] = ws[WS_ST + 12 + lane];
    }
    if (lane < 4) B.base_quat_wxyz[env * 4 + lane] = ws[WS_ST + 4 + lane];
    if (leg) {
        const int o = env * A + 3 * c;
        B.dof_pos[o] = qf4.x; B.dof_pos[o + 1] = qf4.y; B.dof_pos[o + 2] = qf4.z;
        B.dof_vel[o] = vf4.x; B.dof_vel[o + 1] = vf4.y; B.dof_vel[o + 2] = vf4.z;
        B.torques[o] = tf4.x; B.torques[o + 1] = tf4.y; B.torques[o + 2] = tf4.z;
    }
    if (T.i[TI_CAT]) {   // CaT stand-still constraint couples all envs as shipped (go2_cat.py:177-178)
        const bool fast = leg && (fabsf(qd[0]) > 4.0f || fabsf(qd[1]) > 4.0f || fabsf(qd[2]) > 4.0f);
        if (__ballot_sync(B200_FULL_MASK, fast) != 0u && lane == 0) atomicOr(B.global_flags, 1);
    }
    for (int e = lane; e < 3 * L; e += 32) B.link_contact_forces[env * 3 * L + e] = ws[WS_LF + e];
    for (int e = lane; e < 48; e += 32) B.contact_warm[env * 48 + e] = ws[WS_WARM + e];
    const int F = T.i[TI_F];
    if (lane < F) {
        const int l2 = T.i[TI_FEET_LINKS + lane];
        const int b = msi[l2];
        const float *fr = ws + WS_FR + b * 12, *vl = ws + WS_VEL + b * 6, *off = ms + MS_LINK + 3 * l2;
        m33 Rb;
#pragma unroll
        for (int e = 0; e < 9; e++) Rb.m[e] = fr[e];
        const f3 x = mk3(fr[9], fr[10], fr[11]) + mul(Rb, mk3(off[0], off[1], off[2]));
        const f3 v = mk3(vl[3], vl[4], vl[5]) + cross3(mk3(vl[0], vl[1], vl[2]), x);
        float *fp = B.feet_pos + (env * F + lane) * 3, *fv = B.feet_vel + (env * F + lane) * 3;
        fp[0] = p.x + x.x; fp[1] = p.y + x.y; fp[2] = p.z + x.z; fv[0] = v.x; fv[1] = v.y; fv[2] = v.z;
    }
}

// stage the packed robot model into shared memory (once per CTA)
__device__ __forceinline__ void stage_model(const ModelDev &M, const TaskDev &T, float *ms) {
    const int NB = 1 + T.i[TI_A], L = T.i[TI_L], NS = T.i[TI_NSPHERES];
    int *msi = (int *)(ms + MS_INT);
    for (int e = threadIdx.x; e < NB * B200_BODY_STRIDE; e += blockDim.x) ms[MS_BODY + e] = M.body[e];
    for (int e = threadIdx.x; e < L * 3; e += blockDim.x) ms[MS_LINK + e] = M.link_off[e];
    for (int e = threadIdx.x; e < NS * 4; e += blockDim.x) ms[MS_SPH + e] = M.sph[e];
    for (int e = threadIdx.x; e < L; e += blockDim.x) msi[e] = M.link_body[e];
    for (int e = threadIdx.x; e < NS; e += blockDim.x) { msi[B200_MAX_LINKS + e] = M.sph_body[e]; msi[B200_MAX_LINKS + B200_MAX_SPHERES + e] = M.sph_link[e]; }
}

// ragged last CTA: idle warps still take part in the phase barriers of dynamics_warp (1 + 4 per substep, 1 in the last pass)
__device__ __forceinline__ void dynamics_idle_warp(const TaskDev &T) {
#ifndef DYN_NO_PHASE_SYNC
    for (int sub = 0; sub <= T.i[TI_DECIMATION]; sub++) {
        PHASE_SYNC_A();
        if (sub == T.i[TI_DECIMATION]) break;
        PHASE_SYNC_B(); PHASE_SYNC_C(); PHASE_SYNC_D(); PHASE_SYNC_E();
    }
#endif
}

// Which env a warp slot takes.  Inside a CTA the warps wait for each other at the phase barriers, and an SM's four CTAs
// share its issue slots, so a launch is fastest when the envs of a CTA cost the same and every SM gets the same mix.
// dynamics_order_kernel sorts the envs by the solver work they needed two launches ago (contact configurations persist
// over many policy steps), heaviest first: with the block scheduler filling the SMs breadth-first, CTAs 0..147 (the
// heaviest 7 x 148 envs) land one per SM, the next 148 likewise, and so on.  The permutation is delta-encoded
// (slot w runs env w + delta[w]) so that a zeroed buffer is the identity.  Pure scheduling: every env computes exactly
// what it would in any other slot.
#define DYN_COST_BUCKETS 64
#define DYN_COST_PER_BUCKET 48
__global__ void dynamics_order_kernel(const int32_t *cost, int32_t *order_delta, int N, int slots_per_round) {
    __shared__ int start[DYN_COST_BUCKETS];
    for (int b = threadIdx.x; b < DYN_COST_BUCKETS; b += blockDim.x) start[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(&start[min(cost[i] / DYN_COST_PER_BUCKET, DYN_COST_BUCKETS - 1)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {          // exclusive prefix over the buckets, heaviest bucket first
        int acc = 0;
        for (int b = DYN_COST_BUCKETS - 1; b >= 0; b--) { const int n = start[b]; start[b] = acc; acc += n; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int pos = atomicAdd(&start[min(cost[i] / DYN_COST_PER_BUCKET, DYN_COST_BUCKETS - 1)], 1);     // unique rank: a permutation
        // ranks -> warp slots in boustrophedon over the rounds of the block scheduler (one CTA per SM per round): the SM
        // that got the heaviest CTA of a round gets the lightest of the next, so the SMs' total work evens out
        const int round = pos / slots_per_round, r = pos - round * slots_per_round;
        const int len = min(slots_per_round, N - round * slots_per_round);
        const int slot = round * slots_per_round + ((round & 1) ? len - 1 - r : r);
        order_delta[slot] = i - slot;
    }
}

// Host actions, staged: b200_env_step with pinned HOST actions launches this copy kernel right in front of the dynamics
// kernel -- 16 bytes per thread, i.e. 512-byte requests over PCIe instead of the 32-byte sectors per-env reads would make
// -- and the dynamics kernel as its PROGRAMMATIC DEPENDENT (cudaLaunchAttributeProgrammaticStreamSerialization): every copy
// CTA releases the dependent launch first thing, so the dynamics CTAs load their state and run the first forward-kinematics
// pass while the actions cross the bus, and only wait (griddepcontrol.wait = the copy grid has completed and is visible)
// where the first torque needs them.
__global__ void fetch_actions_kernel(const uint4 *__restrict__ src_host, uint4 *__restrict__ dst, int nvec) {
#ifndef B200_WARP_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nvec) dst[i] = src_host[i];
}

template <int C>
__global__ void B200_LAUNCH_BOUNDS(DYN_WARPS_PER_BLOCK * 32, DYN_MIN_BLOCKS)
dynamics_step_kernel(const TaskDev T, const B200Buffers B, const ModelDev M, const TerrainDev tr, const float *actions, const int parity,
                     const int mode) {
    extern __shared__ float smem[];
    float *ms = smem;
#ifndef B200_WARP_EMU
    // a kernel launched as this one's programmatic dependent (the env kernel of a whole host-stepped step) may be placed as
    // soon as every CTA of this grid has got here, i.e. is resident: its CTAs then take the slots this grid's CTAs free
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    stage_model(M, T, ms);
    // the per-step reduction area of the env kernel that follows (episode sums, reset count, level sums): zeroed here so
    // that the step needs no memset node between the kernels
    if (blockIdx.x == 0 && (int)threadIdx.x < T.i[TI_N_SUMS] + 4) B.stats[threadIdx.x] = 0.f;
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#if !defined(B200_WARP_EMU) && !defined(DYN_NO_PIN_LANE)
    // keep lane / warp in registers: without this the compiler re-reads SR_TID (S2R, ~20 cycles) ~260 times per env and
    // policy step to rematerialise them on the address paths of the shared-memory accesses (0.2985 -> 0.2888 ms per step)
    asm volatile("" : "+r"(lane), "+r"(warp));
#endif
    const int slot = blockIdx.x * (blockDim.x >> 5) + warp, N = T.i[TI_NUM_ENVS];
    if (slot >= N) { dynamics_idle_warp(T); return; }
    const bool ordered = parity >= 0 && B.dyn_order != nullptr;
    const int env = ordered ? slot + B.dyn_order[(size_t)parity * N + slot] : slot;
    dynamics_warp<C>(T, B, tr, ms, smem + MS_TOTAL + warp * WS_TOTAL, actions, env, lane,
                     (parity >= 0 && B.dyn_cost) ? B.dyn_cost + (size_t)parity * N + env : nullptr, mode);
}
