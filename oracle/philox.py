"""numpy Philox4x32-10 — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the generator of Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
(SC'11), the counter-based RNG the CUDA sampler and the fused dropout use.  Pinned in
tests/test_oracle.py against the Random123 known-answer vectors.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised over numpy arrays of counters (broadcastable); returns four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)))
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def dropout_keep_mask(n_rows: int, n_cols: int, p: float, seed: int, offset: int) -> np.ndarray:
    """Keep-mask of the fused dropout in K-GEMM's epilogue (noise_gnn_b200/csrc/gemm_simt.cuh::dropout_keep8):
    one philox(counter=(m, c//8, offset_lo, offset_hi), key=(seed_lo, seed_hi)) call serves 8 columns with 16 bits
    each; element (m, c) uses halfword h = c % 8 (word h//2, low half for even h) and is kept iff it is
    >= floor(p * 2^16)."""
    thr = min(max(int(p * 65536.0), 0), 65535)
    cg = (n_cols + 7) // 8
    m = np.arange(n_rows, dtype=np.uint64)[:, None]
    g = np.arange(cg, dtype=np.uint64)[None, :]
    words = philox4x32_10(m, g, offset & 0xFFFFFFFF, (offset >> 32) & 0xFFFFFFFF,
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = np.stack(words, axis=-1)                                   # [rows, groups, 4 words]
    halves = np.stack([w & np.uint32(0xFFFF), w >> np.uint32(16)], axis=-1)   # [rows, groups, 4, (low, high)]
    h = halves.reshape(n_rows, cg * 8)[:, :n_cols]
    return h >= np.uint32(thr)
