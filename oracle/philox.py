"""Philox4x32-10 counter-based random numbers (Salmon et al., SC'11), numpy restatement.

TEST INFRASTRUCTURE.  The product's device-side generator (csrc/philox.cuh) must produce the same
stream: draw(seed, step, env, site, idx) = word (idx & 3) of philox(key=(seed_lo, seed_hi),
counter=(env, step, site, idx >> 2)), mapped to [0,1) as (word >> 8) * 2**-24 (the same 24-bit
mapping torch uses for float32 `torch.rand`).  The reference itself draws from torch's global
generator in data-dependent order, which no fused kernel can reproduce; parity
tests therefore inject these draws into the reference (tools/make_golden.py).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, np.uint32) for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def uniform(seed: int, step: int, env, site: int, idx) -> np.ndarray:
    """float32 uniform in [0,1) for every (env, idx) pair (broadcast)."""
    env = np.asarray(env, np.uint32)
    idx = np.asarray(idx, np.uint32)
    env, idx = np.broadcast_arrays(env, idx)
    w = philox4x32(env, np.uint32(step & 0xFFFFFFFF), np.uint32(site), idx >> np.uint32(2),
                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = idx & np.uint32(3)
    word = np.where(sel == 0, w[0], np.where(sel == 1, w[1], np.where(sel == 2, w[2], w[3])))
    return ((word >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
