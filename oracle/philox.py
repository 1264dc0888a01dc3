"""ORACLE (test infrastructure): numpy restatement of Philox4x32-10 as used for the dropout masks.

Counter = (elem//4 lo, elem//4 hi, site, step), key = (seed lo, seed hi); u = (word >> 8) * 2^-24.
Independent of the CUDA source (`mac_network_b200/csrc/common.cuh`): written from the published
Philox round function (Salmon et al., SC'11) so the in-kernel generator has a second opinion.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox_uniform(seed, site, step, n):
    n4 = (n + 3) // 4
    idx = np.arange(n4, dtype=np.uint64)
    c0 = idx & MASK
    c1 = idx >> np.uint64(32)
    c2 = np.full(n4, site, dtype=np.uint64)
    c3 = np.full(n4, step, dtype=np.uint64)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    words = np.stack([c0, c1, c2, c3], axis=1).reshape(-1)[:n]
    return (words >> np.uint64(8)).astype(np.float64) / 16777216.0
