// tests/warp_emu/cuda_shim.h -- TEST-ONLY synchronous stand-in for the few CUDA runtime calls of
// hcr_genesis_lr_cl_b200/csrc/b200_step.cu, so that the C ABI's host logic (handle state, phase masks, the
// pre-shift / parity protocol, packed copies) runs over the single-warp emulator (emu.h) with host memory.
// Compiled only into tests/warp_emu/libb200step_emu.so; nothing under hcr_genesis_lr_cl_b200/ can load it.
// "Device" memory is host memory, streams and events are no-ops because every launch completes before it returns.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <functional>
#include "emu.h"

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA runtime error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaFuncAttributes { int numRegs; size_t sharedSizeBytes; };

static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 4; return cudaSuccess; }   // "4 SMs": small grids exercise the grid-stride loops
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorUnknown; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
// "device" memory is host memory here: every pointer reads as mapped host memory, so the zero-copy paths of b200_env_step run
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    a->type = cudaMemoryTypeHost; a->device = 0; a->devicePointer = (void *)p; a->hostPointer = (void *)p; return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, F) { a->numRegs = 0; a->sharedSizeBytes = 0; return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }

// run a kernel body (a lambda that calls the __global__ function) on a grid of one-warp blocks
static inline void emu_launch_fn(std::function<void()> fn, int blocks, int blocks_y) {
    static std::function<void()> cur;
    cur = std::move(fn);
    emu_launch([](void *) { cur(); }, nullptr, blocks, blocks_y);
}
