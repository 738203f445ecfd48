"""NumPy restatement of csrc/rls_philox.cuh (Philox4x32-10 counter RNG + the two
integer-exact distributions).  TEST INFRASTRUCTURE: lets the oracle regenerate, bit for
bit, the synthetic matrices / vectors the library generates on the device."""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
UNIFORM01, IH4 = 0, 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0 = c0.astype(np.uint32); c1 = c1.astype(np.uint32); c2 = c2.astype(np.uint32); c3 = c3.astype(np.uint32)
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32); lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32); lo1 = (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_values(seed, idx, stream, component, dist, scale):
    """float32 values for element indices `idx` (uint64 array)."""
    idx = np.asarray(idx, dtype=np.uint64)
    seed = int(seed); stream = int(stream)
    c0 = (idx & _MASK).astype(np.uint32)
    c1 = (idx >> np.uint64(32)).astype(np.uint32)
    c2 = np.full(idx.shape, (stream * 2 + component) & 0xFFFFFFFF, np.uint32)
    c3 = np.full(idx.shape, (stream >> 31) & 0xFFFFFFFF, np.uint32)
    r0, r1, r2, r3 = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    scale = np.float32(scale)
    if dist == UNIFORM01:
        u = (r0 >> np.uint32(8)).astype(np.float32) * np.float32(5.9604644775390625e-08)
        return u * scale
    s = ((r0 >> np.uint32(8)).astype(np.int64) + (r1 >> np.uint32(8)).astype(np.int64) +
         (r2 >> np.uint32(8)).astype(np.int64) + (r3 >> np.uint32(8)).astype(np.int64) - (1 << 25)).astype(np.int32)
    K = np.float32(1.0323827126512697e-07)
    return (s.astype(np.float32) * K) * scale


def philox_vector(dtype, n, seed, stream=0, dist=UNIFORM01, scale=1.0, offset=0):
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    re = philox_values(seed, idx, stream, 0, dist, scale)
    if np.dtype(dtype).kind == "c":
        im = philox_values(seed, idx, stream, 1, dist, scale)
        return (re + 1j * im).astype(np.complex64)
    return re


def philox_matrix(dtype, m, n, seed, dist=IH4, scale=1.0, row_offset=0, m_global=None, chunk_cols=256):
    """Rows [row_offset, row_offset+m) of the global m_global x n matrix, column-major."""
    m_global = m if m_global is None else m_global
    A = np.empty((m, n), dtype=np.dtype(dtype), order="F")
    rows = np.arange(row_offset, row_offset + m, dtype=np.uint64)
    for j0 in range(0, n, chunk_cols):
        j1 = min(n, j0 + chunk_cols)
        cols = np.arange(j0, j1, dtype=np.uint64)
        idx = rows[:, None] + cols[None, :] * np.uint64(m_global)
        re = philox_values(seed, idx, 0, 0, dist, scale)
        if np.dtype(dtype).kind == "c":
            im = philox_values(seed, idx, 0, 1, dist, scale)
            A[:, j0:j1] = re + 1j * im
        else:
            A[:, j0:j1] = re
    return A
