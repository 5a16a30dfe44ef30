"""numpy restatement of the in-kernel RNG of uwu_noise_fwd (uwudiff_b200/csrc/noise.cu).

Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11) with
counter = {group_lo, group_hi, offset_lo, offset_hi}, key = seed.  The reference draws noise with
`torch.randn_like` / `torch.randint` (src/duwu/loss/diffusion.py:68-70,75); matching ATen's thread->counter map
is a non-goal (SURVEY.md §7.2 "RNG parity"), so production-mode RNG is specified HERE and the kernel is tested
against this file: timesteps bit-exact, normals to float tolerance.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
TSTEP_KEY_XOR0, TSTEP_KEY_XOR1 = 0x5851F42D, 0x4C957F2D
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def sample_timesteps(B: int, T: int, seed: int, offset: int) -> np.ndarray:
    b = np.arange(B, dtype=np.uint64)
    z = np.zeros(B, dtype=np.uint64)
    r0, _, _, _ = philox4x32_10(b, z, z + np.uint64(offset & 0xFFFFFFFF), z + np.uint64((offset >> 32) & 0xFFFFFFFF),
                                (seed & 0xFFFFFFFF) ^ TSTEP_KEY_XOR0, ((seed >> 32) & 0xFFFFFFFF) ^ TSTEP_KEY_XOR1)
    return ((r0.astype(np.uint64) * np.uint64(T)) >> np.uint64(32)).astype(np.int64)


def normals(B: int, n_per: int, seed: int, offset: int) -> np.ndarray:
    """N(0,1) noise [B, n_per] exactly as the kernel lays it out (4 consecutive elements per Philox call)."""
    ngroups = (n_per + 3) // 4
    g = (np.arange(B, dtype=np.uint64)[:, None] * np.uint64(ngroups) + np.arange(ngroups, dtype=np.uint64)[None, :]).ravel()
    z = np.zeros_like(g)
    r = philox4x32_10(g & MASK, g >> np.uint64(32), z + np.uint64(offset & 0xFFFFFFFF),
                      z + np.uint64((offset >> 32) & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((g.size, 4), dtype=np.float32)
    for h in range(2):
        u1 = ((r[2 * h] >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0**-24)
        u2 = r[2 * h + 1].astype(np.float32) * np.float32(2.0**-32)
        rad = np.sqrt(np.float32(-2.0) * np.log(u1.astype(np.float64))).astype(np.float32)
        ang = np.float32(6.283185307179586) * u2
        out[:, 2 * h] = rad * np.cos(ang.astype(np.float64)).astype(np.float32)
        out[:, 2 * h + 1] = rad * np.sin(ang.astype(np.float64)).astype(np.float32)
    return out.reshape(B, ngroups * 4)[:, :n_per]
