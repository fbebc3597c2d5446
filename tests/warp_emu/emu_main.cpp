// tests/warp_emu/emu_main.cpp -- TEST-ONLY (see emu.h): C entry points that run the kernels of
// hcr_genesis_lr_cl_b200/csrc on the single-warp emulator with host (numpy) buffers.
#define B200_WARP_EMU 1
#include "emu.h"
#include "../../hcr_genesis_lr_cl_b200/csrc/dynamics_kernel.cuh"
#ifdef EMU_WITH_ENV
#define B200_ENV_DEFINE_KERNELS
#include "../../hcr_genesis_lr_cl_b200/csrc/env_kernel.cuh"
#endif

struct DynArgs { TaskDev T; B200Buffers B; ModelDev M; TerrainDev tr; const float *actions; };

static void dyn_body(void *p) {
    DynArgs *a = (DynArgs *)p;
    if (a->T.i[TI_C] == 4) dynamics_step_kernel<4>(a->T, a->B, a->M, a->tr, a->actions, 0, 0);      // parity 0: zeroed order = identity
    else dynamics_step_kernel<2>(a->T, a->B, a->M, a->tr, a->actions, 0, 0);
}

static ModelDev mk_model(const TaskDev &T, const int *mi, const float *mf) {
    ModelDev M; const int nb = 1 + T.i[TI_A], nl = T.i[TI_L], ns = T.i[TI_NSPHERES];
    M.body = mf; M.link_off = mf + nb * B200_BODY_STRIDE; M.sph = M.link_off + 3 * nl;
    M.link_body = mi + 8; M.sph_body = M.link_body + nl; M.sph_link = M.sph_body + ns;
    return M;
}

extern "C" {
int emu_dynamics_step(const float *tf, const int *ti, const int *mi, const float *mf, const int16_t *hf, int rows, int cols,
                      const B200Buffers *bufs, const float *actions) {
    static DynArgs a;
    memcpy(a.T.f, tf, sizeof a.T.f); memcpy(a.T.i, ti, sizeof a.T.i);
    a.B = *bufs; a.M = mk_model(a.T, mi, mf);
    a.tr.hf = rows > 0 ? hf : nullptr; a.tr.origins = nullptr; a.tr.rows = rows; a.tr.cols = cols; a.tr.levels = a.tr.types = 0;
    a.actions = actions;
    emu_launch(dyn_body, &a, a.T.i[TI_NUM_ENVS]);
    return 0;
}

#ifdef EMU_WITH_ENV
struct EnvArgs { TaskDev T; B200Buffers B; TerrainDev tr; EnvCall call; EnvStageTab tab; };
static int g_env_preset = -1, g_env_specialized = 1;
static void env_body(void *p) {
    EnvArgs *a = (EnvArgs *)p;
    switch (g_env_preset) {            // same selection as launch_env in csrc/b200_step.cu
#define X_CASE(P) case P: env_post_step_kernel_preset<P, false>(a->T, a->B, a->tr, a->call, a->tab); break;
        ENV_FOR_EACH_PRESET(X_CASE)
#undef X_CASE
    default: env_post_step_kernel<false>(a->T, a->B, a->tr, a->call, a->tab);
    }
}
int emu_env_post_step(const float *tf, const int *ti, const int16_t *hf, int rows, int cols, const float *origins, int levels, int types,
                      const B200Buffers *bufs, long long step, float vx_lo, float vx_span, long long hist_step, int phase_mask, int force_reset, int sit_pose, const float *beh8, int gait_cb, int gait_reset) {
    static EnvArgs a;
    memcpy(a.T.f, tf, sizeof a.T.f); memcpy(a.T.i, ti, sizeof a.T.i);
    a.B = *bufs;
    a.tr.hf = rows > 0 ? hf : nullptr; a.tr.origins = origins; a.tr.rows = rows; a.tr.cols = cols; a.tr.levels = levels; a.tr.types = types;
    a.call.step = (uint32_t)step; a.call.vx_lo = vx_lo; a.call.vx_span = vx_span; a.call.hist_step = (uint32_t)hist_step; a.call.phase_mask = phase_mask;
    a.call.force_reset = force_reset; a.call.sit_pose = sit_pose;
    for (int k = 0; k < 8; k++) a.call.beh[k] = beh8 ? beh8[k] : 0.f;
    a.call.gait_cb = gait_cb; a.call.gait_reset = gait_reset;
    a.tab = env_stage_table(a.T, a.B, 1);          // the emulator runs one warp per block
    a.call.finalize = (!force_reset && (phase_mask & PHASE_RESET)) ? 1 : 0;      // like make_call in csrc/b200_step.cu
    a.call.stats_slot = (int)(step % ENV_STATS_RING);
    a.call.inv_episode_length_s = 1.0f / a.T.f[TF_EPISODE_LENGTH_S]; a.call.inv_num_envs = 1.0f / (float)a.T.i[TI_NUM_ENVS];
    a.call.inv_teacher = 0.f; a.call.inv_student = 1.0f / (float)a.T.i[TI_NUM_ENVS];
    g_env_preset = g_env_specialized ? env_match_preset(a.T.i) : -1;
    if (a.T.i[TI_R18] && (phase_mask & PHASE_REWARD) && !force_reset)      // like launch_env in csrc/b200_step.cu
        emu_launch([](void *p) { EnvArgs *e = (EnvArgs *)p; r18_flags_kernel(e->T, e->B, e->call); }, &a, (a.T.i[TI_NUM_ENVS] + R18_FLAGS_BLOCK - 1) / R18_FLAGS_BLOCK);
    emu_launch(env_body, &a, a.T.i[TI_NUM_ENVS]);
    return g_env_preset;
}
void emu_set_env_specialized(int on) { g_env_specialized = on; }

#endif
}
