// b200_step.cu -- C ABI (include/b200_step.h) over the sm_100a kernels.  The product library is built by nvcc only
// (hcr_genesis_lr_cl_b200/build.py); there is no CPU implementation behind these symbols.  With B200_WARP_EMU defined
// (tests/warp_emu, host g++, TEST ONLY) the same host logic runs over the single-warp emulator and a synchronous stand-in
// for the CUDA runtime, so that the host layers above the C ABI can be tested in a container without a GPU.
#ifdef B200_WARP_EMU
#include "cuda_shim.h"
#define B200_ENV_DEFINE_KERNELS
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) emu_launch_fn([&] { kern(__VA_ARGS__); }, (int)dim3(grid).x, (int)dim3(grid).y)
#else
#include <cuda_runtime.h>
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "dynamics_kernel.cuh"
#include "env_kernel.cuh"

static thread_local std::string g_err;
static int fail(const char *what, cudaError_t e = cudaSuccess) {
    g_err = what;
    if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
    return 1;
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(#x, e_); } while (0)

struct B200Handle {
    TaskDev task;
    B200Buffers bufs;
    bool bound = false;
    ModelDev model{};
    TerrainDev terrain{};
    float *d_model_f = nullptr;
    int *d_model_i = nullptr;
    long long launches = 0;
    int dyn_smem = 0, env_smem = 0;
    int sit_pose = 0;
    float beh[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int gait_cb = 0, gait_reset = 0;
    EnvStageTab stage{};
    // side stream of dynamics_order_kernel: forked at an event recorded before the dynamics launch, joined by the next
    // b200_env_post_step
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_enabled = true;
    int env_preset = -1;           // instantiation of the env kernel: index into env_presets.inc, -1 = generic
    float *d_actions = nullptr;    // staging of b200_env_step's host actions (when they cannot be read in place)
    int zero_copy_actions = 2;     // pinned host actions: 2 = staged by a copy kernel the dynamics kernel is the programmatic dependent of (default),
                                   // 1 = read in place by the dynamics kernel, 0 = copy-engine transfer (B200_ZERO_COPY_ACTIONS)
    bool env_pdl = false;          // B200_ENV_PDL=1: env kernel launched as the programmatic dependent of the dynamics kernel (measured: see launch_env)
    bool dyn_just_launched = false; // the last launch on the caller's stream was the dynamics kernel (set by b200_env_step only)
    bool zero_copy_results = true; // the env kernel writes rew | reset | time_out into the caller's pinned slab (B200_ZERO_COPY_RESULTS=0: D2H copy)
    bool stats_zeroed = false;
    bool order_enabled = true;     // dynamics warps take envs sorted by solver cost (dynamics_order_kernel); B200_DYN_ORDER=0: slot w = env w
    long long dyn_launches = 0;    // parity of the cost / order buffers
    int sm_count = 148;
    B200RolloutTargets ro{};       // extra destinations of the next fused post step (b200_set_rollout_targets), cleared by it
    int device = 0;                // the CUDA device the handle was created on (current at b200_create); every entry point runs there
    bool side_pending = false;     // work on the side stream that the next env launch has to join
};

// Every entry point runs on the device the handle was created on, whatever the caller's current device is (the
// reference passes only a device string around, task_registry.py:65,117): a launch on a stream of another device fails.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const B200Handle *h) {
        if (!h) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != h->device && cudaSetDevice(h->device) == cudaSuccess) prev = cur;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};


extern "C" {

const char *b200_last_error(void) { return g_err.c_str(); }

int b200_create(const int32_t *mi, int n_mi, const float *mf, int n_mf, const int32_t *ti, int n_ti, const float *tf, int n_tf,
                B200Handle **out) {
    if (!mi || !mf || !ti || !tf || !out) return fail("b200_create: null argument");
    if (n_ti != TI_COUNT || n_tf != TF_COUNT) return fail("b200_create: task descriptor size does not match this library (header drift)");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return fail("b200_create: no CUDA device");
    B200Handle *h = new B200Handle();
    if (cudaGetDevice(&h->device) != cudaSuccess) { delete h; return fail("b200_create: cudaGetDevice failed"); }
    memcpy(h->task.f, tf, sizeof(float) * TF_COUNT);
    memcpy(h->task.i, ti, sizeof(int) * TI_COUNT);
    const int C = ti[TI_C], A = ti[TI_A], L = ti[TI_L], NS = ti[TI_NSPHERES];
    if (ti[TI_D] != 3 || (C != 2 && C != 4) || A != 3 * C) { delete h; return fail("b200_create: unsupported kinematic tree (need 2 or 4 chains of 3 joints)"); }
    if (L > B200_MAX_LINKS || NS > B200_MAX_SPHERES || ti[TI_F] > B200_MAX_FEET) { delete h; return fail("b200_create: robot exceeds compiled limits"); }
    const int nb = 1 + A;
    if (n_mf != nb * B200_BODY_STRIDE + 3 * L + 4 * NS || n_mi != 8 + L + 2 * NS) { delete h; return fail("b200_create: packed model size mismatch"); }
#define CKH(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { b200_destroy(h); return fail(#x, e_); } } while (0)   /* no leak on failure */
    CKH(cudaMalloc(&h->d_model_f, sizeof(float) * n_mf));
    CKH(cudaMalloc(&h->d_model_i, sizeof(int) * n_mi));
    CKH(cudaMemcpy(h->d_model_f, mf, sizeof(float) * n_mf, cudaMemcpyHostToDevice));
    CKH(cudaMemcpy(h->d_model_i, mi, sizeof(int) * n_mi, cudaMemcpyHostToDevice));
    h->model.body = h->d_model_f; h->model.link_off = h->d_model_f + nb * B200_BODY_STRIDE; h->model.sph = h->model.link_off + 3 * L;
    h->model.link_body = h->d_model_i + 8; h->model.sph_body = h->model.link_body + L; h->model.sph_link = h->model.sph_body + NS;
    h->dyn_smem = dyn_smem_bytes(DYN_WARPS_PER_BLOCK);
    h->env_smem = env_smem_bytes(ENV_WARPS_PER_BLOCK);
    {   // specialised instantiation when the descriptor is one of the built-in presets (B200_ENV_GENERIC=1 forces the generic one)
        const char *g = getenv("B200_ENV_GENERIC");
        h->env_preset = (g && g[0] == '1') ? -1 : env_match_preset(ti);
        CKH(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
        const char *o = getenv("B200_DYN_ORDER");
        h->order_enabled = !(o && o[0] == '0');
        const char *z = getenv("B200_ZERO_COPY_ACTIONS");
        h->zero_copy_actions = (z && z[0] >= '0' && z[0] <= '2') ? z[0] - '0' : 2;
        const char *ep = getenv("B200_ENV_PDL");
        h->env_pdl = ep && ep[0] == '1';
        const char *zr = getenv("B200_ZERO_COPY_RESULTS");
        h->zero_copy_results = !(zr && zr[0] == '0');
    }
    if (h->env_smem > 48 * 1024) {
        CKH(cudaFuncSetAttribute(env_kernel_fn(h->env_preset), cudaFuncAttributeMaxDynamicSharedMemorySize, h->env_smem));
        CKH(cudaFuncSetAttribute(env_kernel_dev_fn(h->env_preset), cudaFuncAttributeMaxDynamicSharedMemorySize, h->env_smem));
    }
    CKH(cudaFuncSetAttribute(dynamics_step_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->dyn_smem));
    CKH(cudaFuncSetAttribute(dynamics_step_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->dyn_smem));
#undef CKH
    *out = h;
    return 0;
}

void b200_destroy(B200Handle *h) {
    if (!h) return;
    DeviceGuard guard(h);
    cudaFree(h->d_model_f); cudaFree(h->d_model_i); cudaFree(h->d_actions);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
}

int b200_set_terrain(B200Handle *h, const int16_t *hf, int rows, int cols, const float *origins, int levels, int types) {
    if (!h) return fail("b200_set_terrain: null handle");
    if (h->task.i[TI_HEIGHTFIELD] && (!hf || rows < 3 || cols < 3)) return fail("b200_set_terrain: heightfield task needs height samples");
    if (h->task.i[TI_TERRAIN_CURRICULUM] && (!origins || levels != h->task.i[TI_NUM_LEVELS] || types != h->task.i[TI_NUM_TYPES]))
        return fail("b200_set_terrain: terrain origins do not match the task descriptor");
    h->terrain.hf = h->task.i[TI_HEIGHTFIELD] ? hf : nullptr; h->terrain.rows = rows; h->terrain.cols = cols;
    h->terrain.origins = origins; h->terrain.levels = levels; h->terrain.types = types;
    return 0;
}

int b200_bind_buffers(B200Handle *h, const B200Buffers *b) {
    if (!h || !b) return fail("b200_bind_buffers: null argument");
    const void *const *p = (const void *const *)b;
    const size_t n = sizeof(B200Buffers) / sizeof(void *);
    for (size_t k = 0; k < n; k++) {
        const bool optional = (&p[k] == (const void *const *)&b->height_cells) || (&p[k] == (const void *const *)&b->dyn_cost) ||
                              (&p[k] == (const void *const *)&b->dyn_order) || (&p[k] == (const void *const *)&b->nonfinite);
        if (!p[k] && !optional) return fail("b200_bind_buffers: null buffer pointer");
    }
    h->bufs = *b; h->bound = true;
    h->stage = env_stage_table(h->task, h->bufs, ENV_WARPS_PER_BLOCK);
    return 0;
}

static int check_ready(B200Handle *h, const char *who) {
    if (!h) return fail("null handle");
    if (!h->bound) { g_err = std::string(who) + ": buffers not bound"; return 1; }
    if (h->task.i[TI_HEIGHTFIELD] && !h->terrain.hf) { g_err = std::string(who) + ": terrain not set"; return 1; }
    return 0;
}

// `host_actions` != NULL: device-visible address of pinned host actions; they are staged into h->d_actions by a copy kernel the
// dynamics kernel is launched as the programmatic dependent of (fetch_actions_kernel), `actions` is ignored
static int launch_dynamics(B200Handle *h, const float *actions, void *stream, int sim_only, const float *host_actions = nullptr) {
    DeviceGuard guard(h);
    const int N = h->task.i[TI_NUM_ENVS];
    const dim3 grid((N + DYN_WARPS_PER_BLOCK - 1) / DYN_WARPS_PER_BLOCK), block(DYN_WARPS_PER_BLOCK * 32);
    cudaStream_t s = (cudaStream_t)stream;
    if (h->task.i[TI_CAT] || h->task.i[TI_R18]) CK(cudaMemsetAsync(h->bufs.global_flags, 0, sizeof(int32_t), s));      // [0] only: [1] ticket, [2] cumulative counter
    // the env -> warp-slot order of the NEXT dynamics launch, from the cost of the PREVIOUS one (the launch enqueued below
    // reads / writes the other halves of dyn_order / dyn_cost): a one-CTA sort on the side stream, under this launch
    const bool ordered = h->order_enabled && h->bufs.dyn_order && h->bufs.dyn_cost;
    const bool reorder = ordered && h->dyn_launches >= 1;
    if (reorder && h->side_enabled) {
        if (!h->side) CK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
        if (!h->ev_fork) CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        if (!h->ev_join) CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        if (h->side_pending) CK(cudaStreamWaitEvent(s, h->ev_join, 0));        // an order kernel nobody joined yet: before this launch reads its half
        CK(cudaEventRecord(h->ev_fork, s));
        CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    }
    const int par = ordered ? (int)(h->dyn_launches & 1) : -1;
    int mode = sim_only ? DYN_MODE_SIM_ONLY : 0;
#ifndef B200_WARP_EMU
    if (host_actions) {
        const int nvec = (int)(((size_t)N * h->task.i[TI_A] * sizeof(float)) / 16);
        if (!h->d_actions) CK(cudaMalloc(&h->d_actions, sizeof(float) * (size_t)N * h->task.i[TI_A]));
        fetch_actions_kernel<<<(nvec + 255) / 256, 256, 0, s>>>((const uint4 *)host_actions, (uint4 *)h->d_actions, nvec);
        h->launches++;
        CK(cudaGetLastError());
        actions = h->d_actions;
        mode |= DYN_MODE_LATE_ACTIONS;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = (size_t)h->dyn_smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (h->task.i[TI_C] == 4) CK(cudaLaunchKernelEx(&cfg, dynamics_step_kernel<4>, h->task, h->bufs, h->model, h->terrain, actions, par, mode));
        else CK(cudaLaunchKernelEx(&cfg, dynamics_step_kernel<2>, h->task, h->bufs, h->model, h->terrain, actions, par, mode));
    } else
#endif
    if (h->task.i[TI_C] == 4) B200_LAUNCH(dynamics_step_kernel<4>, grid, block, h->dyn_smem, s, h->task, h->bufs, h->model, h->terrain, actions, par, mode);
    else B200_LAUNCH(dynamics_step_kernel<2>, grid, block, h->dyn_smem, s, h->task, h->bufs, h->model, h->terrain, actions, par, mode);
    h->dyn_launches++;
    h->launches++;
    h->stats_zeroed = true;
    CK(cudaGetLastError());
    if (reorder) {
        const int next = (int)(h->dyn_launches & 1);                    // parity of the next dynamics launch; its cost half is two launches old
        cudaStream_t run = h->side_enabled ? h->side : s;
        B200_LAUNCH(dynamics_order_kernel, 1, 1024, 0, run, h->bufs.dyn_cost + (size_t)next * N, h->bufs.dyn_order + (size_t)next * N, N,
                    h->sm_count * DYN_WARPS_PER_BLOCK);
        h->launches++;
        CK(cudaGetLastError());
        if (h->side_enabled) { CK(cudaEventRecord(h->ev_join, h->side)); h->side_pending = true; }
    }
    return 0;
}

int b200_dynamics_step(B200Handle *h, const float *actions, void *stream) {
    if (check_ready(h, "b200_dynamics_step")) return 1;
    if (!actions) return fail("b200_dynamics_step: null actions");
    return launch_dynamics(h, actions, stream, 0);
}

int b200_simulator_step(B200Handle *h, const float *actions, void *stream) {
    if (check_ready(h, "b200_simulator_step")) return 1;
    if (!actions) return fail("b200_simulator_step: null actions");
    return launch_dynamics(h, actions, stream, 1);
}

int b200_set_dynamics_order(B200Handle *h, int enabled) {
    if (!h) return fail("b200_set_dynamics_order: null handle");
    h->order_enabled = enabled != 0;
    return 0;
}

int b200_set_side_stream(B200Handle *h, int enabled) {
    if (!h) return fail("b200_set_side_stream: null handle");
    h->side_enabled = enabled != 0;
    return 0;
}

static EnvCall make_call(B200Handle *h, long long step, float lo, float span, long long hist_step, int mask, int force) {
    EnvCall call; call.step = (uint32_t)step; call.vx_lo = lo; call.vx_span = span; call.hist_step = (uint32_t)(hist_step < 0 ? 0 : hist_step); call.phase_mask = mask; call.force_reset = force; call.sit_pose = h->sit_pose;
    for (int k = 0; k < 8; k++) call.beh[k] = h->beh[k];
    call.gait_cb = h->gait_cb; call.gait_reset = h->gait_reset;
    memset(&call.ro, 0, sizeof call.ro);
    if ((mask & PHASE_ALL) == PHASE_ALL && !force) { call.ro = h->ro; memset(&h->ro, 0, sizeof h->ro); }     // one step only
    // extras["episode"] means are finalised by the env kernel's last CTA when the reset phase ran for real
    const int N = h->task.i[TI_NUM_ENVS];
    const int n_teach = max(0, min(N, h->task.i[TI_NUM_TEACHER] - h->task.i[TI_ENV_OFFSET]));   // go2_cts: teacher envs of this rank
    call.finalize = ((mask & PHASE_RESET) && !force) ? 1 : 0;
    call.stats_slot = (int)(step % ENV_STATS_RING);
    call.inv_episode_length_s = 1.0f / h->task.f[TF_EPISODE_LENGTH_S];
    call.inv_num_envs = 1.0f / (float)N;
    call.inv_teacher = h->task.i[TI_NUM_TEACHER] > 0 ? 1.0f / (float)max(n_teach, 1) : 0.f;
    call.inv_student = 1.0f / (float)max(N - n_teach, 1);
    call.dstate = nullptr;
    call.host_slab = nullptr; call.host_slab_vecs = 0;
    return call;
}

static int launch_env(B200Handle *h, long long step, float lo, float span, long long hist_step, int mask, int force, void *stream,
                      const int32_t *dstate = nullptr, void *host_slab = nullptr) {
    DeviceGuard guard(h);
    const int N = h->task.i[TI_NUM_ENVS];
    const int n_sums = h->task.i[TI_N_SUMS];
    cudaStream_t s = (cudaStream_t)stream;
    if ((mask & PHASE_RESET) && !h->stats_zeroed) CK(cudaMemsetAsync(h->bufs.stats, 0, sizeof(float) * (n_sums + 4), s));
    h->stats_zeroed = false;
    EnvCall call = make_call(h, step, lo, span, hist_step, mask, force);
    call.dstate = dstate;
    if (host_slab && call.finalize) { call.host_slab = (uint4 *)host_slab; call.host_slab_vecs = (int)((6 * (size_t)N) / 16); }
    // B200_ENV_PDL=1 (off by default): host-stepped whole steps launch the env kernel as the PROGRAMMATIC DEPENDENT of the
    // dynamics kernel that precedes it on the stream -- its CTAs are placed and set up while the last dynamics CTAs drain, and
    // wait (griddepcontrol.wait at the top of the kernel) for the dynamics grid to complete before they touch global memory.
    // Nothing may sit between the two launches for that, so the side stream (whose ordering kernel only the NEXT dynamics
    // launch reads) is then joined there instead of here.  Measured (go2_ts, 4096 envs, gpurun_out/r2ac_*): steps enqueued
    // back to back 0.1885 -> 0.1875 ms, but an isolated step (cold L2, the bench's `value`) 0.1904 -> 0.1966 ms: not the default.
    const bool pdl = h->env_pdl && h->dyn_just_launched && !dstate && !force && (mask & PHASE_ALL) == PHASE_ALL && !h->task.i[TI_R18];
    h->dyn_just_launched = false;
    if (h->side_pending && !pdl) { CK(cudaStreamWaitEvent(s, h->ev_join, 0)); h->side_pending = false; }      // join the side stream
    if (h->task.i[TI_R18] && (mask & PHASE_REWARD) && !force) {      // bug-compatible mode only: the "any env" bits row 0 follows (R18)
        B200_LAUNCH(r18_flags_kernel, dim3((N + R18_FLAGS_BLOCK - 1) / R18_FLAGS_BLOCK), R18_FLAGS_BLOCK, 0, s, h->task, h->bufs, call);
        h->launches++;
        CK(cudaGetLastError());
    }
    const dim3 grid((N + ENV_WARPS_PER_BLOCK - 1) / ENV_WARPS_PER_BLOCK), block(ENV_WARPS_PER_BLOCK * 32);
#ifndef B200_WARP_EMU
    if (pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = (size_t)h->env_smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, env_kernel_fn(h->env_preset), h->task, h->bufs, h->terrain, call, h->stage));
    } else
#endif
    B200_LAUNCH((dstate ? env_kernel_dev_fn(h->env_preset) : env_kernel_fn(h->env_preset)), grid, block, h->env_smem, s, h->task, h->bufs, h->terrain, call, h->stage);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

int b200_env_post_step(B200Handle *h, long long step, float lo, float span, long long hist_step, int mask, void *stream) {
    if (check_ready(h, "b200_env_post_step")) return 1;
    if ((mask & PHASE_ALL) == 0) return fail("b200_env_post_step: empty phase mask");
    return launch_env(h, step, lo, span, hist_step, mask & PHASE_ALL, 0, stream);
}

int b200_env_step(B200Handle *h, const float *actions, int actions_on_host, long long step, float lo, float span, long long hist_step,
                  float *host_rew, uint8_t *host_reset, uint8_t *host_time_out, void *stream) {
    if (check_ready(h, "b200_env_step")) return 1;
    if (!actions) return fail("b200_env_step: null actions");
    DeviceGuard guard(h);
    const int N = h->task.i[TI_NUM_ENVS], A = h->task.i[TI_A];
    cudaStream_t s = (cudaStream_t)stream;
    const float *dev_actions = actions, *staged = nullptr;
    if (actions_on_host) {
        // Pinned (page-locked, mapped) host memory is addressable from the device, so no copy-engine operation has to sit in
        // front of the launch.  Default (B200_ZERO_COPY_ACTIONS=2): a copy kernel stages the actions with 512-byte requests and
        // the dynamics kernel, its programmatic dependent, waits for them only where the first torque needs them
        // (fetch_actions_kernel).  =1: every env's warp reads its A actions in place (32-byte sectors over PCIe: ~6 us
        // exposed at 4096 envs).  =0, or a buffer the driver does not report as mapped: cudaMemcpyAsync as before.
        const float *mapped = nullptr;
        if (h->zero_copy_actions) {
            cudaPointerAttributes pa;
            if (cudaPointerGetAttributes(&pa, actions) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
                mapped = (const float *)pa.devicePointer;
            else (void)cudaGetLastError();
        }
        if (mapped && h->zero_copy_actions == 2 && ((size_t)N * A * sizeof(float)) % 16 == 0 && ((uintptr_t)mapped & 15) == 0) staged = mapped;
        else if (mapped) dev_actions = mapped;
        else {
            if (!h->d_actions) CK(cudaMalloc(&h->d_actions, sizeof(float) * (size_t)N * A));
            CK(cudaMemcpyAsync(h->d_actions, actions, sizeof(float) * (size_t)N * A, cudaMemcpyHostToDevice, s));
            dev_actions = h->d_actions;
        }
    }
    if (launch_dynamics(h, dev_actions, stream, 0, staged)) return 1;
    // rew | reset | time_out laid out back to back on both sides (B200Simulator allocates them so): one copy ...
    const uint8_t *d_rew = (const uint8_t *)h->bufs.rew_buf;
    const bool packed = host_rew && host_reset && host_time_out && h->bufs.reset_buf == d_rew + 4 * (size_t)N &&
                        h->bufs.time_out_buf == d_rew + 5 * (size_t)N && host_reset == (uint8_t *)host_rew + 4 * (size_t)N &&
                        host_time_out == (uint8_t *)host_rew + 5 * (size_t)N;
    // ... and none at all when the host slab is mapped into the device's address space: the env kernel's finalising CTA
    // writes it over PCIe itself (EnvCall::host_slab; B200_ZERO_COPY_RESULTS=0: copy-engine transfer as before)
    void *slab = nullptr;
    // (one CTA does that copy at the kernel's end: fine for the 24 KB of 4096 envs, slower than the copy engine beyond ~64 KB)
    if (packed && h->zero_copy_results && (6 * (size_t)N) % 16 == 0 && 6 * (size_t)N <= 65536 && ((uintptr_t)d_rew & 15) == 0) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, host_rew) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer &&
            ((uintptr_t)pa.devicePointer & 15) == 0) slab = pa.devicePointer;
        else (void)cudaGetLastError();
    }
    h->dyn_just_launched = true;         // nothing was enqueued on `stream` since the dynamics launch above
    if (launch_env(h, step, lo, span, hist_step, PHASE_ALL, 0, stream, nullptr, slab)) return 1;
    if (slab) return 0;
    if (packed) CK(cudaMemcpyAsync(host_rew, d_rew, 6 * (size_t)N, cudaMemcpyDeviceToHost, s));
    else {
        if (host_rew) CK(cudaMemcpyAsync(host_rew, h->bufs.rew_buf, sizeof(float) * (size_t)N, cudaMemcpyDeviceToHost, s));
        if (host_reset) CK(cudaMemcpyAsync(host_reset, h->bufs.reset_buf, (size_t)N, cudaMemcpyDeviceToHost, s));
        if (host_time_out) CK(cudaMemcpyAsync(host_time_out, h->bufs.time_out_buf, (size_t)N, cudaMemcpyDeviceToHost, s));
    }
    return 0;
}

int b200_env_step_device(B200Handle *h, const float *actions, int32_t *ss, double sit_init_percent, void *stream) {
    if (check_ready(h, "b200_env_step_device")) return 1;
    if (!actions || !ss) return fail("b200_env_step_device: null argument");
    if (h->task.i[TI_R18]) return fail("b200_env_step_device: the bug-compatible R18 mode needs host-stepped calls");
    DeviceGuard guard(h);
    // host: coin < p, coin = k * 2^-24  <=>  k < ceil(p * 2^24)
    double thr = sit_init_percent > 0.0 ? sit_init_percent * 16777216.0 : 0.0;
    uint32_t thr_q = (uint32_t)thr; if ((double)thr_q < thr) thr_q++;
    if (thr_q > 16777216u) thr_q = 16777216u;
    B200_LAUNCH(step_advance_kernel, 1, 32, 0, (cudaStream_t)stream, ss, (uint32_t)h->task.i[TI_SEED_LO], (uint32_t)h->task.i[TI_SEED_HI], thr_q,
                h->task.i[TI_BEHAVIOR], (int)SITE_HOST);
    h->launches++;
    CK(cudaGetLastError());
    if (launch_dynamics(h, actions, stream, 0)) return 1;
    return launch_env(h, 0, 0.f, 0.f, 0, PHASE_ALL, 0, stream, ss);
}

int b200_set_rollout_targets(B200Handle *h, const B200RolloutTargets *t) {
    if (!h) return fail("b200_set_rollout_targets: null handle");
    if (!t) { memset(&h->ro, 0, sizeof h->ro); return 0; }
    if ((t->ep_return != nullptr) != (t->ep_length != nullptr) || (t->ep_return != nullptr) != (t->ep_stats != nullptr))
        return fail("b200_set_rollout_targets: ep_return, ep_length and ep_stats go together");
    h->ro = *t;
    return 0;
}

int b200_set_step_flags(B200Handle *h, int sit_pose) {
    if (!h) return fail("b200_set_step_flags: null handle");
    h->sit_pose = sit_pose ? 1 : 0;
    return 0;
}

int b200_set_behavior(B200Handle *h, const float *ranges8, int gait_callback, int gait_reset) {
    if (!h || !ranges8) return fail("b200_set_behavior: null argument");
    if (gait_callback < 0 || gait_callback >= B200_MAX_GAITS || gait_reset < 0 || gait_reset >= B200_MAX_GAITS) return fail("b200_set_behavior: gait index out of range");
    // stored as (lower, span) with span rounded like torch_rand_float's python-float subtraction
    for (int k = 0; k < 4; k++) { h->beh[2 * k] = ranges8[2 * k]; h->beh[2 * k + 1] = (float)((double)ranges8[2 * k + 1] - (double)ranges8[2 * k]); }
    h->gait_cb = gait_callback; h->gait_reset = gait_reset;
    return 0;
}

int b200_reset_all(B200Handle *h, long long step, float lo, float span, void *stream) {
    if (check_ready(h, "b200_reset_all")) return 1;
    return launch_env(h, step, lo, span, 0, PHASE_RESET, 1, stream);
}

int b200_kernel_info(B200Handle *h, const char *kernel, int *regs, int *smem, int *blocks_per_sm, int *block_threads) {
    if (!h || !kernel) return fail("b200_kernel_info: null argument");
    DeviceGuard guard(h);
    cudaFuncAttributes fa; int nb = 0, threads = 0, dyn = 0;
    if (!strcmp(kernel, "dynamics")) {
        threads = DYN_WARPS_PER_BLOCK * 32; dyn = h->dyn_smem;
        if (h->task.i[TI_C] == 4) { CK(cudaFuncGetAttributes(&fa, dynamics_step_kernel<4>)); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, dynamics_step_kernel<4>, threads, dyn)); }
        else { CK(cudaFuncGetAttributes(&fa, dynamics_step_kernel<2>)); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, dynamics_step_kernel<2>, threads, dyn)); }
    } else if (!strcmp(kernel, "env")) {
        threads = ENV_WARPS_PER_BLOCK * 32; dyn = h->env_smem;
        CK(cudaFuncGetAttributes(&fa, env_kernel_fn(h->env_preset))); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, env_kernel_fn(h->env_preset), threads, dyn));
    } else return fail("b200_kernel_info: unknown kernel name");
    if (regs) *regs = fa.numRegs;
    if (smem) *smem = (int)fa.sharedSizeBytes + dyn;
    if (blocks_per_sm) *blocks_per_sm = nb;
    if (block_threads) *block_threads = threads;
    return 0;
}

const char *b200_env_kernel_variant(B200Handle *h) { return (h && h->env_preset >= 0) ? kEnvPresetName[h->env_preset] : "generic"; }

long long b200_launch_count(B200Handle *h) { return h ? h->launches : 0; }

int b200_stats_ring(void) { return ENV_STATS_RING; }

int b200_device(B200Handle *h) { return h ? h->device : -1; }

#ifdef DYN_TIMING
/* diagnosis build only (-DDYN_TIMING, tools/probe_dyn_timing.py): per-env {start ns, end ns, sweeps, sweep-rows} of the last launch */
int b200_debug_dyn_timing(unsigned long long *out, int n_envs) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out, g_dyn_timing, sizeof(unsigned long long) * 4 * (size_t)n_envs));
    return 0;
}
#endif

}  // extern "C"

#ifdef B200_WARP_EMU
#include "env_kernels_host.cu"
#include "env_kernels_dev.cu"
#endif
