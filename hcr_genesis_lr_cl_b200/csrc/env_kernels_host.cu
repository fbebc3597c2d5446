// env_kernels_host.cu -- the host-stepped instantiations of the fused post_physics_step kernel and the small kernels
// around it (r18_flags_kernel, step_advance_kernel), in a translation unit of their own: compiled beside b200_step.cu and
// with -fmad=false (see env_kernel.cuh).  Under the test emulator (B200_WARP_EMU, one g++ translation unit, contraction
// off) the file is included at the end of b200_step.cu.
#ifndef B200_WARP_EMU
#define B200_ENV_DEFINE_KERNELS
#include <cuda_runtime.h>
#include "env_kernel.cuh"
#endif

EnvKernelFn env_kernel_fn(int preset) {
    switch (preset) {
#define X_CASE(P) case P: return env_post_step_kernel_preset<P, false>;
        ENV_FOR_EACH_PRESET(X_CASE)
#undef X_CASE
    default: return env_post_step_kernel<false>;
    }
}
