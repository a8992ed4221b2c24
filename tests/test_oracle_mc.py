"""Marching-cubes oracle: PARITY UNPINNED (VTK absent, the reference has no golden mesh).
What can be checked without VTK: the mesh is a closed, consistently oriented 2-manifold
with outward normals on closed inputs (all 256 cases incl. ambiguous faces via noise), the
vertices sit on the iso-crossing edges, and the geometry frame follows
surface_process.py:100-161 / converters.py:55-63 (padding offsets, z of the piece, Y flip)
and the reference test's bounds envelope (tests/test_mesh_generation.py:62-83)."""
import numpy as np
import pytest


def manifold_stats(V, T):
    e = np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]])
    key = e[:, 0] * (len(V) + 1) + e[:, 1]
    rkey = e[:, 1] * (len(V) + 1) + e[:, 0]
    edge_manifold = len(np.unique(key)) == len(key)
    closed = np.array_equal(np.sort(key), np.sort(rkey))
    p = V[T].astype(np.float64)
    vol = np.einsum("ij,ij->i", p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() / 6
    area = np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1).sum() / 2
    chi = len(V) - len(key) // 2 + len(T)
    return edge_manifold, closed, vol, area, chi


def sphere_u8(n=40, r=12.3):
    z, y, x = np.mgrid[:n, :n, :n]
    d = np.sqrt((x - (n - 1) / 2) ** 2 + (y - (n - 1) / 2) ** 2 + (z - (n - 1) / 2) ** 2)
    return np.clip((r - d) * 40 + 127, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("flip", [False, True])
def test_sphere_is_closed_oriented_manifold(orc, flip):
    r = 12.3
    V, T = orc.marching_cubes(sphere_u8(r=r), 127, flip_y=flip)
    em, closed, vol, area, chi = manifold_stats(V, T)
    assert em and closed and chi == 2
    # outward normals -> positive signed volume in the output frame, flipped or not
    assert vol == pytest.approx(4 / 3 * np.pi * r ** 3, rel=0.02)
    assert area == pytest.approx(4 * np.pi * r * r, rel=0.02)
    assert (V[:, 1] <= 0).all() if flip else (V[:, 1] >= 0).all()


def test_all_cases_closed_manifold_on_noise(orc):
    rng = np.random.default_rng(0)
    seen = set()
    for trial in range(6):
        vol = np.zeros((14, 15, 16), np.uint8)
        vol[1:-1, 1:-1, 1:-1] = rng.integers(0, 2, (12, 13, 14)) * 255
        b = vol > 127
        c = sum(b[(k >> 2) & 1:b.shape[0] - 1 + ((k >> 2) & 1), (k >> 1) & 1:b.shape[1] - 1 + ((k >> 1) & 1),
                  (k & 1):b.shape[2] - 1 + (k & 1)].astype(int) << k for k in range(8))
        seen |= set(np.unique(c).tolist())
        V, T = orc.marching_cubes(vol, 127, flip_y=bool(trial & 1))
        em, closed, volm, _, _ = manifold_stats(V, T)
        assert em and closed and volm > 0
    assert len(seen) == 256


def test_vertices_lie_on_crossing_edges_and_interpolate(orc):
    rng = np.random.default_rng(1)
    vol = rng.integers(-500, 1500, (6, 7, 8)).astype(np.int16)
    iso = 226.0
    sp = (0.5, 0.75, 1.5)
    V, T = orc.marching_cubes(vol, iso, spacing=sp, origin_index=(-1, -1, 3), flip_y=True)
    idx = np.empty_like(V, dtype=np.float64)
    idx[:, 0] = V[:, 0] / sp[0] + 1
    idx[:, 1] = -V[:, 1] / sp[1] + 1
    idx[:, 2] = V[:, 2] / sp[2] - 3
    frac = idx - np.floor(idx + 1e-6)
    assert ((frac > 1e-4).sum(axis=1) <= 1).all()           # on a grid edge
    assert idx.min() >= -1e-4 and (idx.max(axis=0) <= np.array([7, 6, 5]) + 1e-4).all()
    # spot check one vertex per axis against the formula
    inside = vol >= iso
    p = np.argwhere(inside[:, :, :-1] != inside[:, :, 1:])[0]
    s0, s1 = float(vol[tuple(p)]), float(vol[p[0], p[1], p[2] + 1])
    t = np.float32(np.float32(iso - s0) / np.float32(s1 - s0))
    want = np.array([np.float32(np.float32(p[2] - 1) + t) * np.float32(sp[0]),
                     -(np.float32(p[1] - 1) * np.float32(sp[1])), np.float32(p[0] + 3) * np.float32(sp[2])], np.float32)
    assert (np.abs(V - want).max(axis=1) == 0).any()
    assert T.min() == 0 and T.max() == len(V) - 1 and len(np.unique(T)) == len(V)


def test_cube_bounds_like_reference_test(orc):
    """tests/test_mesh_generation.py:22-83: a solid cube contoured at iso 0.5-ish stays
    within a margin of its voxel bounds."""
    vol = np.zeros((20, 20, 20), np.uint8)
    vol[5:15, 5:15, 5:15] = 255
    V, T = orc.marching_cubes(vol, 127, flip_y=False)
    assert len(V) and len(T)
    assert V.min() >= 4 and V.max() <= 15
    em, closed, volm, _, chi = manifold_stats(V, T)
    assert em and closed and chi == 2 and volm == pytest.approx(10.0 ** 3, rel=0.05)


def test_degenerate_inputs(orc):
    V, T = orc.marching_cubes(np.zeros((3, 3, 3), np.uint8), 127)
    assert V.shape == (0, 3) and T.shape == (0, 3)
    V, T = orc.marching_cubes(np.full((3, 3, 3), 255, np.uint8), 127)
    assert V.shape == (0, 3) and T.shape == (0, 3)
    V, T = orc.marching_cubes(np.array([[[0, 255]]], np.uint8), 127)  # one edge, no cell
    assert V.shape == (1, 3) and T.shape == (0, 3)
