"""Host-side scalar draws from the same Philox4x32-10 stream the kernels use (csrc/philox.cuh).

Used for the per-step scalars the reference draws once per batch on the host, e.g. the sit-pose coin of
tron1_pf_ee.py:204: u = philox(key = seed, counter = (0xffffffff, step, SITE_HOST, idx >> 2))[idx & 3].
"""
M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF


def philox4x32(c, k):
    c0, c1, c2, c3 = c
    k0, k1 = k
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def host_uniform(seed: int, step: int, site: int, idx: int = 0) -> float:
    w = philox4x32((0xFFFFFFFF, step & MASK, site, idx >> 2), (seed & 0x7FFFFFFF, (seed >> 32) & 0x7FFFFFFF))
    return (w[idx & 3] >> 8) * 2.0 ** -24


def uniform_block(seed: int, step: int, env_ids, site: int, cols):
    """The kernels' draw `u(seed, step, env, site, idx)` (csrc/philox.cuh) for every (env, idx) pair: float32 [len(env_ids), len(cols)].

    Plugin mode draws the simulator-side randomisation (domain randomisation on reset, pushes, terrain-curriculum
    levels: genesis_simulator.py:140-158,665-739) from this stream with the env's global id and the policy-step counter,
    i.e. exactly the numbers the fused env kernel draws for the same env, step and site."""
    import numpy as np
    env = np.asarray(env_ids, np.uint64).reshape(-1, 1)
    idx = np.asarray(cols, np.uint64).reshape(1, -1)
    env, idx = np.broadcast_arrays(env, idx)
    c0, c1 = env.astype(np.uint64), np.full(env.shape, step & MASK, np.uint64)
    c2, c3 = np.full(env.shape, site, np.uint64), idx >> np.uint64(2)
    k0, k1 = np.uint64(seed & 0x7FFFFFFF), np.uint64((seed >> 32) & 0x7FFFFFFF)
    m = np.uint64(MASK)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2           # < 2^64: operands are 32-bit values
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m, p1 & m, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m, p0 & m
        k0, k1 = (k0 + np.uint64(W0)) & m, (k1 + np.uint64(W1)) & m
    sel = idx & np.uint64(3)
    word = np.where(sel == 0, c0, np.where(sel == 1, c1, np.where(sel == 2, c2, c3)))
    return ((word >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
