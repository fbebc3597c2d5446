// cuda_compat.cuh -- the kernels are plain CUDA C++ for sm_100a.  The only indirection is this header:
// when B200_WARP_EMU is defined (tests/warp_emu, host g++), the CUDA built-ins used by the kernels are
// provided by a single-warp lock-step emulator so that kernel logic can be checked against the oracle
// in a container without a GPU.  The product library (libb200step.so) is always built by nvcc with
// B200_WARP_EMU undefined; there is no CPU path in it.
#pragma once
#ifdef B200_WARP_EMU
#include "emu.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define B200_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#endif

#define B200_FULL_MASK 0xffffffffu

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float comp3(f3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

// row-major 3x3
struct m33 { float m[9]; };
__device__ __forceinline__ f3 mul(const m33 &A, f3 b) {
    return mk3(A.m[0] * b.x + A.m[1] * b.y + A.m[2] * b.z, A.m[3] * b.x + A.m[4] * b.y + A.m[5] * b.z,
               A.m[6] * b.x + A.m[7] * b.y + A.m[8] * b.z);
}
__device__ __forceinline__ m33 mul(const m33 &A, const m33 &B) {
    m33 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
    return C;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(B200_FULL_MASK, v, o);
    return v;
}
