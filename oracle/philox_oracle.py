"""Philox4x32-10 + Box-Muller oracle (numpy) -- TEST INFRASTRUCTURE ONLY.

The reference draws eps with torch's global generator
(bayeformers/nn/parameters/gaussian.py:100); the B200 build replaces that by
a counter-based stream so that backward can regenerate eps instead of storing
it and so that data-parallel replicas agree without communication (SURVEY.md
section 8e).  There is no reference implementation of this stream, so this
file states the contract the CUDA kernels implement (DESIGN.md "Philox
contract") and is pinned by the published Random123 known-answer vectors for
philox4x32-10 (Salmon et al., SC'11; kat_vectors in the Random123
distribution), checked in tests/test_oracle_golden.py.

Contract
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (quad, sample_id, tensor_id, step)     quad = flat_element_index >> 2
    out     = philox4x32_10(counter, key) -> (r0, r1, r2, r3)
    element 4*quad+0, +1  <- box_muller(r0, r1) -> (radius*cos, radius*sin)
    element 4*quad+2, +3  <- box_muller(r2, r3)
    box_muller(a, b): u = a * 2^-32 + 2^-33           in (0, 1]
                      f = b * 2^-31 + (2^-32 - 1)     in (-1, 1]
                      radius = sqrt(-2 ln u), theta = pi * f
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised philox4x32 with 10 rounds.  Inputs are broadcastable uint32
    arrays (or python ints); returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0  # 64-bit products of 32-bit operands: no overflow
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def box_muller(a: np.ndarray, b: np.ndarray, dtype=np.float64):
    """Two uniforms (uint32) -> two N(0,1) variates, per the contract above.
    dtype=float64 gives the mathematically exact value of the contract; the
    device evaluates it in fp32 with MUFU approximations (abs err ~1e-6)."""
    a = a.astype(dtype)
    b = b.astype(dtype)
    u = a * dtype(2.0 ** -32) + dtype(2.0 ** -33)
    f = b * dtype(2.0 ** -31) + dtype(2.0 ** -32 - 1.0)
    radius = np.sqrt(dtype(-2.0) * np.log(u))
    theta = dtype(np.pi) * f
    return radius * np.cos(theta), radius * np.sin(theta)


def philox_normal(n: int, seed: int, step: int, tensor_id: int, sample_id: int, dtype=np.float64) -> np.ndarray:
    """eps[0:n] of (tensor_id, sample_id, step) under `seed`."""
    nq = (n + 3) // 4
    quad = np.arange(nq, dtype=np.uint64)
    r0, r1, r2, r3 = philox4x32_10(quad, sample_id, tensor_id, step, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    n0, n1 = box_muller(r0, r1, dtype)
    n2, n3 = box_muller(r2, r3, dtype)
    out = np.stack([n0, n1, n2, n3], axis=1).reshape(-1)
    return out[:n]


# Random123 known-answer vectors for philox4x32-10: (counter, key, expected)
KAT_PHILOX4X32_10 = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF),
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def dropout_keep_mask(n: int, p: float, seed: int, step: int, site_id: int) -> np.ndarray:
    """Keep mask (uint8, 1 = kept) of the fused dropout + residual + LayerNorm kernels
    (include/bayeformers_b200.h, bf_resln_fwd / bf_dropout_mask).  The reference's
    examples use torch's nn.Dropout inside the HF host model; this states the
    counter-based replacement so that backward regenerates the mask instead of storing it.

    Contract
        counter = (k & 0xffffffff, k >> 32, 0x80000000 | site_id, step)    k = flat_element_index >> 3
        out     = philox4x32_10(counter, key = seed) -> (r0, r1, r2, r3)
        u16 of element 8k + 2i     = r_i & 0xffff
        u16 of element 8k + 2i + 1 = r_i >> 16
        keep    = u16 >= min(floor(p * 65536 + 0.5), 65535);  p <= 0 keeps everything
    """
    if p <= 0:
        return np.ones(n, dtype=np.uint8)
    thr = min(int(np.floor(np.float32(p).astype(np.float64) * 65536.0 + 0.5)), 65535)
    no = (n + 7) // 8
    k = np.arange(no, dtype=np.uint64)
    r = philox4x32_10(k & MASK32, k >> np.uint64(32), 0x80000000 | (site_id & 0x7FFFFFFF), step,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = np.empty((no, 8), dtype=np.uint32)
    for i in range(4):
        u[:, 2 * i] = r[i] & np.uint32(0xFFFF)
        u[:, 2 * i + 1] = r[i] >> np.uint32(16)
    return (u.reshape(-1)[:n] >= thr).astype(np.uint8)


def attention_keep_mask(B: int, H: int, T: int, p: float, seed: int, step: int, site_id: int) -> np.ndarray:
    """Keep mask (uint8 [B, H, T, T], 1 = kept) of the attention kernels (include/bayeformers_b200.h, bf_attention_fwd /
    bf_attention_dropout_mask; both kernel families -- tcgen05 and mma.sync -- apply this one function).  The reference
    leaves attention dropout to torch's generator inside the host model; this states the counter-based replacement.

    Contract  (row = (b * H + h) * T + q, key k)
        counter = (row, (k % 8) // 2 + 4 * (k // 32), 0x40000000 | site_id, step)
        out     = philox4x32_10(counter, key = seed) -> (r0, r1, r2, r3)
        u16     = r_{(k // 8) % 4} & 0xffff  for even k,  >> 16  for odd k
        keep    = u16 >= min(round(p * 65536), 65535);  p <= 0 keeps everything
    """
    if p <= 0:
        return np.ones((B, H, T, T), dtype=np.uint8)
    thr = min(int(np.rint(np.float32(p).astype(np.float64) * 65536.0)), 65535)
    rows = np.arange(B * H * T, dtype=np.uint64)[:, None]
    k = np.arange(T, dtype=np.uint64)[None, :]
    c1 = (k % np.uint64(8)) // np.uint64(2) + np.uint64(4) * (k // np.uint64(32))
    r = philox4x32_10(np.broadcast_to(rows, (B * H * T, T)).copy(), np.broadcast_to(c1, (B * H * T, T)).copy(),
                      0x40000000 | (site_id & 0x3FFFFFFF), step, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    word = ((k // np.uint64(8)) % np.uint64(4)).astype(np.int64)
    words = np.stack([np.broadcast_to(x, (B * H * T, T)) for x in r], axis=0)  # [4, rows, T]
    sel = np.take_along_axis(words, np.broadcast_to(word, (B * H * T, T))[None], axis=0)[0]
    u16 = np.where((k % np.uint64(2)) == 0, sel & np.uint32(0xFFFF), sel >> np.uint32(16))
    return (u16 >= thr).astype(np.uint8).reshape(B, H, T, T)
