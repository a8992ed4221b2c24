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


# ---- the 256-case table itself, checked against first principles (no generator, no product code) ----
def _parse_mc_tables():
    """Reads the numbers out of csrc/mc_tables.h (the table both the CUDA kernels and the CPU checker
    compile in); everything it is checked against below is derived here from the cube's geometry."""
    import re
    from pathlib import Path
    txt = (Path(__file__).resolve().parents[1] / "invesalius3_b200" / "csrc" / "mc_tables.h").read_text()
    ntri = [int(v) for v in re.search(r"B2V_MC_NTRI\[256\] = \{(.*?)\};", txt, re.S).group(1).replace("\n", " ").split(",") if v.strip()]
    body = re.search(r"B2V_MC_TRI\[256\]\[15\] = \{(.*?)\n\};", txt, re.S).group(1)
    tri = [[int(v) for v in row.split(",") if v.strip()] for row in re.findall(r"\{([^{}]*)\}", body)]
    emask = [int(v, 0) for v in re.search(r"B2V_MC_EDGEMASK\[256\] = \{(.*?)\};", txt, re.S).group(1).replace("\n", " ").split(",") if v.strip()]
    assert len(ntri) == 256 and len(tri) == 256 and all(len(r) == 15 for r in tri) and len(emask) == 256
    return ntri, tri, emask


def _cube_edges():
    """edge id = axis * 4 + cu + 2 * cv (cu, cv: offsets along the two other axes in increasing axis
    order); corner bit = x + 2 y + 4 z. Returns id -> (corner a, corner b, midpoint)."""
    edges = {}
    for axis in range(3):
        u, v = [a for a in range(3) if a != axis]
        for cu in (0, 1):
            for cv in (0, 1):
                p0 = [0, 0, 0]; p0[u] = cu; p0[v] = cv
                p1 = list(p0); p1[axis] = 1
                bit = lambda p: p[0] + 2 * p[1] + 4 * p[2]   # noqa: E731
                edges[axis * 4 + cu + 2 * cv] = (bit(p0), bit(p1), (np.array(p0) + np.array(p1)) / 2.0)
    return edges


def test_table_cases_from_first_principles():
    """Every one of the 256 cases: the triangles use exactly the crossing edges; inside the cell the
    patch is an oriented manifold (interior edges shared by two triangles, opposite directions); its
    boundary lies on the cube's faces and, face by face, is what the states of the face's four corners
    dictate — one segment between the two crossing edges, or, on an ambiguous face, two segments that
    each cut off an INSIDE corner. That rule depends on the face's corners only, so two cells that
    share a face draw the same segments on it: the surface is watertight across cells. All segments
    wind the same way around the inside."""
    ntri, tri, emask = _parse_mc_tables()
    edges = _cube_edges()
    corner_xyz = {c: np.array([c & 1, (c >> 1) & 1, (c >> 2) & 1], float) for c in range(8)}
    faces = []   # (axis, side, outward normal, its 4 corners, its 4 edges)
    for axis in range(3):
        for side in (0, 1):
            n = np.zeros(3); n[axis] = 1.0 if side else -1.0
            cs = [c for c in range(8) if corner_xyz[c][axis] == side]
            es = [e for e, (a, b, _) in edges.items() if a in cs and b in cs]
            faces.append((n, cs, es))
    winding = set()
    for case in range(256):
        inside = {c: bool((case >> c) & 1) for c in range(8)}
        crossing = {e for e, (a, b, _) in edges.items() if inside[a] != inside[b]}
        t = tri[case]
        assert all(v == -1 for v in t[3 * ntri[case]:]) and all(0 <= v < 12 for v in t[:3 * ntri[case]]), case
        tris = [tuple(t[3 * k: 3 * k + 3]) for k in range(ntri[case])]
        assert all(len(set(tr)) == 3 for tr in tris), case
        assert {e for tr in tris for e in tr} == crossing, case
        assert emask[case] == sum(1 << e for e in crossing), case
        directed = {}
        for tr in tris:
            for a, b in ((tr[0], tr[1]), (tr[1], tr[2]), (tr[2], tr[0])):
                assert (a, b) not in directed, case       # no edge twice in the same direction
                directed[(a, b)] = True
        boundary = [(a, b) for (a, b) in directed if (b, a) not in directed]
        # every interior edge is matched by its reverse (checked by construction above); the boundary:
        seen = set()
        for n, cs, es in faces:
            fc = [e for e in es if e in crossing]
            segs = [(a, b) for (a, b) in boundary if a in es and b in es]
            seen.update(segs)
            assert len(fc) in (0, 2, 4), case
            assert len(segs) == len(fc) // 2, (case, fc, segs)
            assert sorted(e for s in segs for e in s) == sorted(fc), case
            for a, b in segs:
                if len(fc) == 2:
                    ins = [c for c in cs if inside[c]]
                else:          # ambiguous face: the segment joins the two face edges that meet in an inside corner
                    shared = set(edges[a][:2]) & set(edges[b][:2])
                    assert len(shared) == 1 and inside[next(iter(shared))], (case, a, b)
                    ins = list(shared)
                m = (edges[a][2] + edges[b][2]) / 2.0
                side_vec = np.cross(n, edges[b][2] - edges[a][2])
                signs = {np.sign(np.dot(side_vec, corner_xyz[c] - m)) for c in ins}
                assert len(signs) == 1 and 0.0 not in signs, (case, a, b)
                winding.add(next(iter(signs)))
        assert len(seen) == len(boundary), case         # every boundary segment lies in a face of the cube
    assert len(winding) == 1                            # one winding convention for all 256 cases


# ---- envelope from the reference's own project: samples/Cranium.inv3 ---------------------------------
def mesh_volume_area(V, F):
    a, b, c = (V[F[:, k]].astype(np.float64) for k in range(3))
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0), float(0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1).sum())


def cranium_mask(cranium, i):
    shape = tuple(int(v) for v in cranium["full_shape"])
    m = np.unpackbits(cranium[f"mask_{i}_bits_full"])[: int(np.prod(shape))].reshape(shape) * np.uint8(255)
    assert int((m == 255).sum()) == int(cranium[f"mask_{i}_count_full"])
    pad = np.zeros(tuple(s + 2 for s in shape), np.uint8)       # fill_border_holes (surface_process.py:52-68)
    pad[1:-1, 1:-1, 1:-1] = m
    return pad


@pytest.mark.parametrize("i", [0, 1])
def test_cranium_surface_envelope(orc, cranium, i):
    """The only surface numbers the reference ships: surface_N.plist of samples/Cranium.inv3 records
    the volume of the mesh InVesalius built from mask_N (657 705.6 and 3 161 711.5 mm^3) — after
    its smoothing and decimation, so an envelope rather than a golden mesh. Contouring the reference's
    own mask (iso 127, padded like create_surface_piece, y flipped) must enclose that volume to
    within 2 % (measured: -0.8 % and -0.5 %; the voxel count times the voxel volume lies in between)
    and be a closed, outward-oriented surface."""
    sx, sy, sz = (float(v) for v in cranium["spacing"])
    V, F = orc.marching_cubes(cranium_mask(cranium, i), 127, (sx, sy, sz), (-1, -1, -1), True)
    vol, area = mesh_volume_area(V, F)
    want = float(cranium[f"surface_{i}_volume_mm3"])
    voxels = int(cranium[f"mask_{i}_count_full"]) * sx * sy * sz
    assert abs(vol - want) / want < 0.02, (vol, want)
    assert abs(vol - voxels) / voxels < 0.01, (vol, voxels)
    assert vol > 0 and area > 0
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    key = e[:, 0].astype(np.int64) * (len(V) + 1) + e[:, 1]
    rev = e[:, 1].astype(np.int64) * (len(V) + 1) + e[:, 0]
    assert len(np.unique(key)) == len(key) and np.array_equal(np.sort(key), np.sort(rev))   # closed, consistently oriented
