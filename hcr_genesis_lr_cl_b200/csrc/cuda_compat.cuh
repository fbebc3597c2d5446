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

// sin / cos of a bounded angle on the special-function unit (2 MUFU instead of the ~40-instruction sincosf)
__device__ __forceinline__ void fast_sincosf(float x, float *s, float *c) {
#ifdef B200_WARP_EMU
    *s = sinf(x); *c = cosf(x);
#else
    __sincosf(x, s, c);
#endif
}

// sum over the first C lanes (C = 2 or 4), result in every lane: two butterfly rounds inside the aligned 4-lane group,
// then one broadcast -- the chain quantities of the dynamics kernel live in lanes 0..C-1 only.
template <int C>
__device__ __forceinline__ float lead_sum(float v) {
    if (C == 4) v += __shfl_xor_sync(B200_FULL_MASK, v, 2);
    v += __shfl_xor_sync(B200_FULL_MASK, v, 1);
    return __shfl_sync(B200_FULL_MASK, v, 0);
}

// max over the warp of a non-negative float: IEEE ordering of non-negative floats is the ordering of their bit patterns,
// so one integer REDUX does it (SASS REDUX.MAX.U32) instead of five shuffle + max rounds.
__device__ __forceinline__ float warp_max_nonneg(float v) {
#ifdef B200_WARP_EMU
    unsigned u = (unsigned)__float_as_int(v);       // the same integer max on the bit patterns as REDUX (NaN patterns included: warp-uniform)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned w = __shfl_xor_sync(B200_FULL_MASK, u, o); u = w > u ? w : u; }
    return __int_as_float((int)u);
#else
    return __uint_as_float(__reduce_max_sync(B200_FULL_MASK, __float_as_uint(v)));
#endif
}

// 1 / sqrt(x) for x well inside the normal range: the bare MUFU.RSQ (rsqrtf() wraps it in a denormal fix-up, 4 instructions)
__device__ __forceinline__ float fast_rsqrtf(float x) {
#ifdef B200_WARP_EMU
    return 1.0f / sqrtf(x);
#else
    float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#endif
}

// ---- explicit shared-memory accesses by 32-bit shared-window address: the inner loop of the contact solver keeps ONE
// address register and immediate offsets, instead of generic pointers that the compiler re-derives from SR_TID /
// SR_CgaCtaId under register pressure (S2R on the address path of every access).
#ifdef B200_WARP_EMU
typedef uintptr_t smaddr_t;
__device__ __forceinline__ smaddr_t sm_addr(const void *p) { return (smaddr_t)p; }
__device__ __forceinline__ float4 lds128(smaddr_t a) { return *(const float4 *)a; }
__device__ __forceinline__ float2 lds64(smaddr_t a) { return *(const float2 *)a; }
__device__ __forceinline__ float lds32(smaddr_t a) { return *(const float *)a; }
__device__ __forceinline__ void sts128(smaddr_t a, float x, float y, float z, float w) { float *p = (float *)a; p[0] = x; p[1] = y; p[2] = z; p[3] = w; }
#else
typedef uint32_t smaddr_t;
__device__ __forceinline__ smaddr_t sm_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(smaddr_t a) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ float2 lds64(smaddr_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float lds32(smaddr_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts128(smaddr_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
#endif

// ---- 4-byte asynchronous global -> shared copy (LDGSTS): the loaded word never passes through a register, so a slow source
// (pinned host memory read over PCIe) stalls nobody until cp_async_wait_all().  Emulator: a plain copy.
#ifdef B200_WARP_EMU
__device__ __forceinline__ void cp_async_f32(float *dst_smem, const float *src_gmem) { *dst_smem = *src_gmem; }
__device__ __forceinline__ void cp_async_wait_all() {}
#else
__device__ __forceinline__ void cp_async_f32(float *dst_smem, const float *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

// ---- programmatic dependent launch: wait until the grid this one depends on has completed and its writes are visible
// (a no-op for a launch without a programmatic dependency)
__device__ __forceinline__ void grid_dependency_wait() {
#ifndef B200_WARP_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// ---- TMA (bulk async copy) + mbarrier helpers: 1-D cp.async.bulk between global and shared memory (sm_90+; SASS UBLKCP).
// Sizes and both addresses must be multiples of 16 bytes.  Under the test emulator they degrade to memcpy.
#ifdef B200_WARP_EMU
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { (void)bar; (void)count; }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { (void)bar; (void)bytes; }
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) { (void)bar; memcpy(dst_smem, src_gmem, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) { (void)bar; (void)phase; emu_sync(); }   // all lanes' copies are issued before the wait
__device__ __forceinline__ void tma_store_1d(void *dst_gmem, const void *src_smem, uint32_t bytes) { memcpy(dst_gmem, src_smem, bytes); }
__device__ __forceinline__ void tma_store_commit_wait() {}
__device__ __forceinline__ void fence_proxy_async() {}
#define B200_TMA_SIZE_OK(bytes) true
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    for (;;) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        if (ok) break;
        __nanosleep(200);                                  // leave the issue slots to the CTAs that already have their data
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_1d(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#define B200_TMA_SIZE_OK(bytes) (((bytes) & 15u) == 0u)
#endif
