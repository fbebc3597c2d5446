// philox.cuh -- Philox4x32-10 (Salmon et al., SC'11) counter-based draws for the fused environment kernel.
// Keyed so that results do not depend on launch geometry or GPU count:
//   u(seed, step, env, site, idx) = word (idx & 3) of philox(key=(seed_lo, seed_hi), ctr=(env, step, site, idx>>2))
// mapped to [0,1) with the 24-bit rule torch uses for float32.  Replaces the reference's global-generator
// call sites torch_rand_float / torch.rand_like / gs.rand (math_utils.py:78-81, genesis_simulator.py:669-739).
#pragma once
#include "cuda_compat.cuh"

struct Philox4 { uint32_t w[4]; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 p; p.w[0] = c0; p.w[1] = c1; p.w[2] = c2; p.w[3] = c3; return p;
}

struct EnvRng {
    uint32_t k0, k1, env, step;
    __device__ __forceinline__ Philox4 block(int site, int blk) const { return philox4x32_10(env, step, (uint32_t)site, (uint32_t)blk, k0, k1); }
    __device__ __forceinline__ float u(int site, int idx) const {
        Philox4 p = block(site, idx >> 2);
        const int s = idx & 3;
        const uint32_t w = s == 0 ? p.w[0] : (s == 1 ? p.w[1] : (s == 2 ? p.w[2] : p.w[3]));
        return (float)(w >> 8) * 5.9604644775390625e-08f;
    }
};

// (upper - lower) * u + lower with torch's fp32 rounding (no FMA contraction)
__device__ __forceinline__ float rand_range(float lo, float span, float u) { return __fadd_rn(__fmul_rn(span, u), lo); }
