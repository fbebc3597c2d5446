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
