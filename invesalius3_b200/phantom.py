"""Seeded synthetic CT phantom: int16 [dz][dy][dx].

air -1000; soft-tissue ellipsoid (0.42 of the dims) +40; skull shell (outer 0.40, inner
0.36) +1200; 8 bone spheres (+700, radius 0.06*min dim); trabecular texture
300*sin(.35x)sin(.31y)sin(.27z) inside the shell interior; Gaussian noise sigma 25;
clipped to [-1024, 3071]. Generated slab by slab so 1024^3 fits in host memory.
"""
from __future__ import annotations

import numpy as np


def ct(shape, seed: int = 0, noise: float = 25.0, zrange=None) -> np.ndarray:
    """The whole volume, or only its planes zrange = (z0, z1) (identical values either way:
    the noise of slice z depends on (seed, z) alone, so shards can be generated in place)."""
    dz, dy, dx = (int(s) for s in shape)
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0.3, 0.7, size=(8, 3)) * np.array([dz, dy, dx])
    rad = 0.06 * min(dz, dy, dx)
    if zrange is not None:
        return _ct_planes(dz, dy, dx, seed, noise, centres, rad, int(zrange[0]), int(zrange[1]))
    return _ct_planes(dz, dy, dx, seed, noise, centres, rad, 0, dz)


def _ct_planes(dz, dy, dx, seed, noise, centres, rad, zlo, zhi) -> np.ndarray:
    out = np.empty((zhi - zlo, dy, dx), dtype=np.int16)
    y = np.arange(dy, dtype=np.float32)[None, :, None]
    x = np.arange(dx, dtype=np.float32)[None, None, :]
    cz, cy, cx = (dz - 1) / 2.0, (dy - 1) / 2.0, (dx - 1) / 2.0
    ey = ((y - cy) / max(dy, 1)) ** 2
    ex = ((x - cx) / max(dx, 1)) ** 2
    step = max(1, (1 << 22) // max(1, dy * dx))
    for z0 in range(zlo, zhi, step):
        z1 = min(zhi, z0 + step)
        z = np.arange(z0, z1, dtype=np.float32)[:, None, None]
        r2 = ((z - cz) / max(dz, 1)) ** 2 + ey + ex  # normalised radius^2 (1.0 = full dim)
        v = np.full(r2.shape, -1000.0, dtype=np.float32)
        v[r2 <= 0.42 ** 2] = 40.0
        shell = (r2 <= 0.40 ** 2) & (r2 >= 0.36 ** 2)
        v[shell] = 1200.0
        inner = r2 < 0.36 ** 2
        tex = 300.0 * np.sin(0.35 * x) * np.sin(0.31 * y) * np.sin(0.27 * z)
        v = np.where(inner, v + tex, v)
        for c in centres:
            d2 = (z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2
            v[np.broadcast_to(d2 <= rad * rad, v.shape) & inner] = 700.0
        if noise > 0:
            for k, zz in enumerate(range(z0, z1)):
                v[k] += np.random.default_rng([seed, zz]).normal(0.0, noise, size=(dy, dx)).astype(np.float32)
        np.clip(v, -1024, 3071, out=v)
        out[z0 - zlo:z1 - zlo] = np.rint(v).astype(np.int16)
    return out


def first_seed_in_range(volume: np.ndarray, z: int, lo: int, hi: int):
    """(x, y, z) of the first voxel (raveled order) of slice z with lo <= v <= hi."""
    sl = volume[z]
    idx = np.flatnonzero((sl >= lo) & (sl <= hi))
    if idx.size == 0:
        raise ValueError("no in-range voxel in that slice")
    yy, xx = divmod(int(idx[0]), sl.shape[1])
    return (xx, yy, z)
