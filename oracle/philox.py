"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11), integer-exact restatement.

The reference samples with `torch.multinomial` (HF `_sample`), whose CUDA RNG stream cannot be reproduced
by a custom kernel.  Oracle and kernel therefore share this counter-based generator:
    u(seed, row, frame, group) = (philox4x32_10(key=(seed_lo, seed_hi), ctr=(row, frame, group, 0))[0] >> 8) * 2^-24
and sample by inverse CDF over the post-filter probabilities (oracle/sampler.py).
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(key, ctr):
    """key: (k0,k1) ints, ctr: (c0,c1,c2,c3) ints -> 4 uint32 outputs as python ints."""
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    c0, c1, c2, c3 = [int(c) & 0xFFFFFFFF for c in ctr]
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        hi0, lo0 = (p0 >> 32) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        hi1, lo1 = (p1 >> 32) & 0xFFFFFFFF, p1 & 0xFFFFFFFF
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniform(seed: int, row: int, frame: int, group: int) -> float:
    """Uniform in [0,1) with 24 random bits; exactly representable in fp32."""
    out = philox4x32_10((seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF), (row, frame, group, 0))
    return float(np.float32((out[0] >> 8) * (1.0 / 16777216.0)))
