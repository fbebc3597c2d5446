// env_kernels_dev.cu -- the device-stepped instantiations of the fused post_physics_step kernel (EnvCall::dstate, see
// env_kernel.cuh and b200_env_step_device in include/b200_step.h), in a translation unit of their own so that nvcc compiles
// them beside b200_step.cu instead of after it, with -fmad=false like the host-stepped ones.  Under the test emulator (B200_WARP_EMU, one g++ translation unit) the file
// is included at the end of b200_step.cu.
#ifndef B200_WARP_EMU
#include <cuda_runtime.h>
#include "env_kernel.cuh"
#endif

EnvKernelFn env_kernel_dev_fn(int preset) {
    switch (preset) {
#define X_CASE(P) case P: return env_post_step_kernel_preset<P, true>;
        ENV_FOR_EACH_PRESET(X_CASE)
#undef X_CASE
    default: return env_post_step_kernel<true>;
    }
}
