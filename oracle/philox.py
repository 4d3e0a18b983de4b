"""Philox4x32-10 + Box-Muller in numpy: the stream definition of the in-kernel eps draws.

TEST INFRASTRUCTURE ONLY.  The reference draws eps with ``torch.randn_like`` from the
global generator (cVAE.py:1130-1133); production kernels cannot reproduce that stream,
so per-step parity injects eps and the production stream is *defined* here:

    key     = (seed_lo, seed_hi)
    counter = (group, step_lo, step_hi, stream)        group = element_index // 4
    4 x u32 -> u = ((w >> 9) + 0.5) * 2^-23  (exact in fp32, strictly inside (0,1))
    (n0, n1) = sqrt(-2 ln u0) * (cos, sin)(2 pi u1);  (n2, n3) likewise from (u2, u3)

element_index = row * Z + col inside one minibatch; ``stream`` 0 = training eps,
1 = test-time eps (cVAE.py:1207).  Philox4x32-10 itself is the published Random123
algorithm (Salmon et al., SC'11), pinned by its known-answer vectors.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: [...,4] uint32, key: [...,2] uint32 -> [...,4] uint32."""
    c = [np.asarray(counter[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            if r < 9:
                k0 = (k0 + W0).astype(np.uint32)
                k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def uniforms(seed: int, step: int, n_elems: int, stream: int = 2):
    """u = ((w >> 9) + 0.5) * 2^-23 of element e = word (e & 3) of counter group e >> 2: the dropout stream (stream 2) of the
    end-to-end classifier -- unit (row b, column j of the concatenated hidden layers) has e = b * sum(widths) + j and is
    KEPT when u >= p."""
    groups = (n_elems + 3) // 4
    ctr = np.zeros((groups, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(groups, dtype=np.uint32)
    ctr[:, 1] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, 2] = np.uint32((step >> 32) & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32(stream)
    key = np.zeros((groups, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    w = philox4x32_10(ctr, key)
    u = ((w >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
    return u.reshape(-1)[:n_elems]


def normals(seed: int, step: int, n_elems: int, stream: int = 0):
    """The first n_elems eps values of minibatch `step` for Philox key `seed` (float32)."""
    groups = (n_elems + 3) // 4
    ctr = np.zeros((groups, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(groups, dtype=np.uint32)
    ctr[:, 1] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, 2] = np.uint32((step >> 32) & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32(stream)
    key = np.zeros((groups, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    w = philox4x32_10(ctr, key)
    u = ((w >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
    out = np.empty((groups, 4), dtype=np.float32)
    two_pi = np.float32(6.283185307179586)
    for a in (0, 2):
        r = np.sqrt(np.float32(-2.0) * np.log(u[:, a]))
        th = two_pi * u[:, a + 1]
        out[:, a] = r * np.cos(th)
        out[:, a + 1] = r * np.sin(th)
    return out.reshape(-1)[:n_elems]
