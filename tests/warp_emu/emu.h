// tests/warp_emu/emu.h -- TEST-ONLY single-warp lock-step emulator of the CUDA built-ins the kernels use.
//
// Purpose: check the logic of hcr_genesis_lr_cl_b200/csrc/*.cuh against the oracle in the build container
// (which has no GPU) before spending GPU time.  It is compiled only into tests/warp_emu/libwarpemu.so, which
// nothing under hcr_genesis_lr_cl_b200/ can load; the product has no CPU path.  Each of the 32 lanes of a warp
// is a ucontext coroutine; every warp collective (__shfl*, __ballot, __syncwarp) is a rendezvous, so a missing
// __syncwarp() between a shared-memory write and a read by another lane shows up deterministically.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <algorithm>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__
#define B200_LAUNCH_BOUNDS(t, b)

struct EmuDim { int x, y, z; };
extern EmuDim emu_thread_idx[32];
extern EmuDim emu_block_idx, emu_block_dim, emu_grid_dim;
extern int emu_cur_lane;
#define threadIdx (emu_thread_idx[emu_cur_lane])
#define blockIdx emu_block_idx
#define blockDim emu_block_dim
#define gridDim emu_grid_dim

uint32_t emu_exchange(uint32_t v, int src_lane);
uint32_t emu_ballot(bool pred);
void emu_sync();

static inline uint32_t emu_bits(float v) { uint32_t u; memcpy(&u, &v, 4); return u; }
static inline float emu_float(uint32_t u) { float v; memcpy(&v, &u, 4); return v; }

static inline float __shfl_sync(unsigned, float v, int src) { return emu_float(emu_exchange(emu_bits(v), src & 31)); }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)emu_exchange((uint32_t)v, src & 31); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return emu_exchange(v, src & 31); }
static inline float __shfl_xor_sync(unsigned, float v, int m) { return emu_float(emu_exchange(emu_bits(v), emu_cur_lane ^ m)); }
static inline int __shfl_xor_sync(unsigned, int v, int m) { return (int)emu_exchange((uint32_t)v, emu_cur_lane ^ m); }
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int m) { return emu_exchange(v, emu_cur_lane ^ m); }
static inline float __shfl_down_sync(unsigned, float v, int d) { return emu_float(emu_exchange(emu_bits(v), emu_cur_lane + d < 32 ? emu_cur_lane + d : emu_cur_lane)); }
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline unsigned __ballot_sync(unsigned, bool p) { return emu_ballot(p); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_sync(); }
static inline void __syncthreads() { emu_sync(); }   // the emulator runs one warp per block
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
static inline void __threadfence() {}
#define __restrict__
static inline int __float_as_int(float v) { return (int)emu_bits(v); }
static inline float __int_as_float(int v) { return emu_float((uint32_t)v); }
// volatile keeps g++ from contracting these into FMAs, like the _rn intrinsics on the device
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float atomicAdd(float *p, float v) { float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }
using std::max;
using std::min;

typedef void (*EmuKernelBody)(void *args);
// run `body` on a grid of `blocks` x `blocks_y` blocks of one warp each
void emu_launch(EmuKernelBody body, void *args, int blocks, int blocks_y = 1);
