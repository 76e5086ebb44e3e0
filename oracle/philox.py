"""TEST INFRASTRUCTURE ONLY — pure-Python Philox4x32-10 and the scheduled-sampling mask rule.

Philox4x32-10 is the counter-based generator of Salmon et al., "Parallel Random Numbers: As
Easy as 1, 2, 3" (SC'11); the known-answer vectors in tests/test_oracle_cpu.py are the ones
published with Random123 (kat_vectors: philox4x32 10).

The reference draws its scheduled-sampling mask as ``torch.rand(length) < epoch/max_epochs``
(``mr_gen/model/lstm_with_sampling/lstm_with_sample.py:389``) from the global, unseeded CPU
generator: one Bernoulli per timestep shared by the whole batch.  BASELINE.json's north_star
replaces that with a documented counter-based draw per (timestep, sample):

    counter = (lo32(offset + t), hi32(offset + t), b, 0)        key = (lo32(seed), hi32(seed))
    r       = philox4x32_10(counter, key)[0]
    u       = float32(r >> 8) * 2**-24                            (24-bit uniform in [0, 1))
    mask[t, b] = u < float32(prob)

``shared=True`` uses b = 0 for every sample, which reproduces the reference's
"one decision per timestep" shape.
"""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = (int(v) & MASK32 for v in counter)
    k0, k1 = (int(v) & MASK32 for v in key)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK32
        hi1, lo1 = p1 >> 32, p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK32, lo1, (hi0 ^ c3 ^ k1) & MASK32, lo0
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return c0, c1, c2, c3


def sampling_mask(seed: int, offset: int, prob: float, T: int, B: int, shared: bool = False):
    """bool mask [T, B] — True means "feed the model's own prediction back"."""
    out = np.zeros((T, B), dtype=bool)
    p = np.float32(prob)
    key = (seed & MASK32, (seed >> 32) & MASK32)
    for t in range(T):
        pos = offset + t
        for b in range(B):
            r = philox4x32_10((pos & MASK32, (pos >> 32) & MASK32, 0 if shared else b, 0), key)[0]
            u = np.float32(r >> 8) * np.float32(2.0 ** -24)
            out[t, b] = u < p
    return out
