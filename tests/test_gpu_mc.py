"""GPU marching cubes vs the canonical CPU checker: triangle indices bit-exact, vertex
coordinates within 1e-5 (they are in fact compared for exact equality)."""
import numpy as np
import pytest
from scipy import ndimage

pytestmark = pytest.mark.gpu

from test_oracle_mc import manifold_stats, sphere_u8  # noqa: E402


@pytest.fixture(scope="module")
def sp():
    from invesalius3_b200 import device, surface_process
    device.require_cuda()
    return surface_process


def _check(sp, orc, vol, iso, spacing=(1, 1, 1), z0=0, flip=True, padding=(0, 0, 0)):
    V, F = sp.contour(vol, iso, spacing, z0, flip, padding)
    oi = (-padding[0], -padding[1], z0 - padding[2])
    Vo, Fo = orc.marching_cubes(vol, iso, spacing, oi, flip)
    assert V.shape == Vo.shape and F.shape == Fo.shape, (V.shape, Vo.shape, F.shape, Fo.shape)
    assert np.array_equal(F, Fo)
    assert np.abs(V - Vo).max(initial=0) <= 1e-5
    assert np.array_equal(V, Vo)
    return V, F


@pytest.mark.parametrize("shape", [(2, 2, 2), (3, 5, 7), (7, 9, 33), (9, 10, 64), (12, 17, 97), (20, 33, 130)])
def test_mc_noise_u8(sp, orc, shape):
    rng = np.random.default_rng(sum(shape))
    vol = (rng.integers(0, 2, shape) * 255).astype(np.uint8)
    _check(sp, orc, vol, 127, flip=True)
    _check(sp, orc, vol, 127, spacing=(0.9570312, 0.9570312, 1.5), z0=20, flip=False, padding=(1, 1, 0))


@pytest.mark.parametrize("shape", [(5, 6, 8), (16, 24, 40), (21, 64, 96)])
def test_mc_smooth_i16(sp, orc, shape):
    rng = np.random.default_rng(3)
    f = ndimage.gaussian_filter(rng.normal(size=shape), 1.5)
    vol = (f / np.abs(f).max() * 2000).astype(np.int16)
    for iso in (226, -142, 0.5, 3071):
        _check(sp, orc, vol, iso, spacing=(0.5, 0.75, 1.25))


def test_mc_mask_values_and_markers(sp, orc):
    """Binary path: masks hold 0/1/2/253/254/255 (mask.py value code); inside <=> >= 127."""
    rng = np.random.default_rng(5)
    vol = rng.choice(np.array([0, 1, 2, 253, 254, 255], np.uint8), size=(10, 12, 50))
    _check(sp, orc, vol, 127)


def test_mc_sphere_manifold(sp, orc):
    V, F = _check(sp, orc, sphere_u8(48, 17.2), 127)
    em, closed, vol, area, chi = manifold_stats(V, F)
    assert em and closed and chi == 2 and vol == pytest.approx(4 / 3 * np.pi * 17.2 ** 3, rel=0.02)


def test_mc_empty_and_full(sp):
    for v in (0, 255):
        V, F = sp.contour(np.full((4, 5, 6), v, np.uint8), 127)
        assert V.shape == (0, 3) and F.shape == (0, 3)
    V, F = sp.contour(np.array([[[0, 255]]], np.uint8), 127)
    assert V.shape == (1, 3) and F.shape == (0, 3)
    with pytest.raises(TypeError):
        sp.contour(np.zeros((4, 5, 6), np.float32), 127)


def test_contour_piece_matches_reference_pipeline(sp, orc):
    """create_surface_piece semantics (surface_process.py:97-186): ROI on the padded mask,
    pad_image with 0 (mask) / iinfo.min (image), padding offsets, z offset roi.start."""
    rng = np.random.default_rng(9)
    dz, dy, dx = 45, 20, 37
    body = (ndimage.gaussian_filter(rng.normal(size=(dz, dy, dx)), 1.3) > 0).astype(np.uint8) * 255
    body[0] = 255  # touches the volume border: fill_border_holes closes it
    mm = np.zeros((dz + 1, dy + 1, dx + 1), np.uint8)
    mm[1:, 1:, 1:] = body
    img = (ndimage.gaussian_filter(rng.normal(size=(dz, dy, dx)), 1.3) * 3000).astype(np.int16)
    spacing = (0.9570312, 0.9570312, 1.5)
    piece_size, o_piece = 20, 1
    n_pieces = int(round(dz / piece_size + 0.5))            # surface.py:1365
    for i in range(n_pieces):
        roi = slice(i * piece_size, (i + 1) * piece_size + o_piece)   # surface.py:1375-1381
        for fbh in (True, False):
            # --- binary path
            a = np.array(mm[roi.start + 1:roi.stop + 1, 1:, 1:])
            pb, pt = roi.start == 0, roi.stop >= dz
            if fbh:
                pad = np.zeros((a.shape[0] + pb + pt, dy + 2, dx + 2), np.uint8)
                pad[int(pb):int(pb) + a.shape[0], 1:-1, 1:-1] = a
                a, padding = pad, (1, 1, int(pb))
            else:
                padding = (0, 0, 0)
            V, F = sp.contour_piece(None, mm, roi, spacing, from_binary=True, fill_border_holes=fbh)
            if a.shape[0] == 0:
                assert len(V) == 0
                continue
            Vo, Fo = orc.marching_cubes(a, 127, spacing, (-padding[0], -padding[1], roi.start - padding[2]), True)
            assert np.array_equal(F, Fo) and np.array_equal(V, Vo)
            # --- default path: two iso values on the int16 image
            b = np.array(img[roi])
            if fbh:
                pad = np.full((b.shape[0] + pb + pt, dy + 2, dx + 2), np.iinfo(np.int16).min, np.int16)
                pad[int(pb):int(pb) + b.shape[0], 1:-1, 1:-1] = b
                b = pad
            V, F = sp.contour_piece(img, None, roi, spacing, 226, 3071, from_binary=False, fill_border_holes=fbh)
            V1, F1 = orc.marching_cubes(b, 226, spacing, (-padding[0], -padding[1], roi.start - padding[2]), True)
            V2, F2 = orc.marching_cubes(b, 3071, spacing, (-padding[0], -padding[1], roi.start - padding[2]), True)
            assert np.array_equal(V, np.concatenate([V1, V2]))
            assert np.array_equal(F, np.concatenate([F1, F2 + len(V1)]))


def _piece_worker(args):
    """Runs in a SPAWNED process, like the reference's surface workers (surface.py:1368-1369)."""
    from invesalius3_b200 import surface_process
    return surface_process.create_surface_piece(*args)


def test_create_surface_piece_in_spawned_workers(sp, orc, tmp_path):
    """The reference's own entry (20 arguments in, .vtp file name out; surface_process.py:71-201) over
    the piece loop of SurfaceManager.AddNewActor (surface.py:1360-1381), each piece in a spawned
    process: the files hold exactly the meshes of contour_piece / the oracle."""
    import multiprocessing as mp
    rng = np.random.default_rng(12)
    dz, dy, dx = 45, 24, 40
    body = (ndimage.gaussian_filter(rng.normal(size=(dz, dy, dx)), 1.3) > 0).astype(np.uint8) * 255
    img = (ndimage.gaussian_filter(rng.normal(size=(dz, dy, dx)), 1.3) * 3000).astype(np.int16)
    img_fn, mask_fn = str(tmp_path / "matrix.dat"), str(tmp_path / "mask.dat")
    np.memmap(img_fn, mode="w+", dtype=np.int16, shape=img.shape)[:] = img
    mm = np.memmap(mask_fn, mode="w+", dtype=np.uint8, shape=(dz + 1, dy + 1, dx + 1))
    mm[:] = 0
    mm[1:, 1:, 1:] = body
    mm.flush()
    spacing = (0.9570312, 0.9570312, 1.5)
    piece_size, o_piece = 20, 1
    n_pieces = int(round(dz / piece_size + 0.5))
    rois = [slice(i * piece_size, (i + 1) * piece_size + o_piece) for i in range(n_pieces)]
    jobs = []
    for from_binary in (True, False):
        for roi in rois:
            jobs.append((img_fn, img.shape, "int16", mask_fn, (dz + 1, dy + 1, dx + 1), "uint8", roi, spacing, "CONTOUR", 226,
                         3071, 0.0, 0.0, 0, "en", False, from_binary, "Default", 0, True))
    with mp.get_context("spawn").Pool(2) as pool:
        names = pool.map(_piece_worker, jobs)
    for job, fn in zip(jobs, names):
        roi, from_binary = job[6], job[16]
        assert fn.endswith("_%d_%d.vtp" % (roi.start, roi.stop))
        V, F = sp.read_vtp(fn)
        Vw, Fw = sp.contour_piece(img, mm, roi, spacing, 226, 3071, from_binary=from_binary, fill_border_holes=True,
                                  index_dtype=np.int64)
        assert np.array_equal(V, Vw) and np.array_equal(F, Fw), (roi, from_binary)
        import os
        os.unlink(fn)
    with pytest.raises(NotImplementedError):
        sp.create_surface_piece(*(jobs[-1][:17] + ("InVesalius 3.b2",) + jobs[-1][18:]))


def test_mc_512_properties(sp):
    """BASELINE config-2 size, properties only: closed oriented manifold (the padded phantom
    mask does not touch the border), counts consistent, every index used."""
    import torch
    from invesalius3_b200 import device as dev, phantom
    from invesalius3_b200.mesh import marching_cubes
    vol = phantom.ct((192, 512, 512), seed=2)
    t = torch.from_numpy(vol).cuda()
    mask = dev.threshold(t, 226, 3071)
    mask[0] = 0; mask[-1] = 0; mask[:, 0] = 0; mask[:, -1] = 0; mask[:, :, 0] = 0; mask[:, :, -1] = 0
    V, F = marching_cubes(mask, 127, (1, 1, 1), (0, 0, 0), True)
    assert V.shape[0] > 10000 and F.shape[0] > 10000
    F64 = F.to(torch.int64)
    assert int(F64.min()) == 0 and int(F64.max()) == V.shape[0] - 1
    e = torch.cat([F64[:, [0, 1]], F64[:, [1, 2]], F64[:, [2, 0]]])
    key = e[:, 0] * (V.shape[0] + 1) + e[:, 1]
    rkey = e[:, 1] * (V.shape[0] + 1) + e[:, 0]
    assert torch.unique(key).numel() == key.numel()
    assert torch.equal(torch.sort(key).values, torch.sort(rkey).values)
    assert bool((V[:, 1] <= 0).all()) and bool(torch.isfinite(V).all())


def test_mc_wide_rows(sp, orc):
    """2048-wide rows (BASELINE config 5 geometry), uint8 mask and int16 image."""
    rng = np.random.default_rng(21)
    shape = (4, 19, 2048)
    f = ndimage.gaussian_filter(rng.normal(size=shape), (1, 2, 5))
    _check(sp, orc, (f > 0).astype(np.uint8) * 255, 127, spacing=(1, 1, 1))
    _check(sp, orc, (f / np.abs(f).max() * 3000).astype(np.int16), 226, spacing=(0.5, 0.5, 2.0), z0=100)


def test_mc_config5_shard_properties(sp):
    """One GPU's share of BASELINE config 5 (2048 x 2048 x 1024 over 8 GPUs = 128 slices):
    threshold -> marching cubes on 2^29 voxels; closed oriented manifold, indices in range."""
    import torch
    from invesalius3_b200 import device as dev
    from invesalius3_b200.mesh import marching_cubes
    nz, ny, nx = 128, 2048, 2048
    z = torch.arange(nz, device="cuda", dtype=torch.float32)[:, None, None]
    y = torch.arange(ny, device="cuda", dtype=torch.float32)[None, :, None]
    x = torch.arange(nx, device="cuda", dtype=torch.float32)[None, None, :]
    img = (1500.0 * torch.sin(0.011 * x) * torch.sin(0.013 * y) * torch.cos(0.05 * z)).to(torch.int16)
    del x, y, z
    mask = dev.threshold(img, 226, 3071)
    del img
    mask[0] = 0; mask[-1] = 0; mask[:, 0] = 0; mask[:, -1] = 0; mask[:, :, 0] = 0; mask[:, :, -1] = 0
    V, F = marching_cubes(mask, 127, (1, 1, 1), (0, 0, 0), True)
    nv, nt = V.shape[0], F.shape[0]
    assert nv > 10 ** 6 and nt > 2 * 10 ** 6
    assert int(F.min()) == 0 and int(F.max()) == nv - 1
    F64 = F.to(torch.int64)
    del F
    e = torch.cat([F64[:, [0, 1]], F64[:, [1, 2]], F64[:, [2, 0]]])
    key = e[:, 0] * (nv + 1) + e[:, 1]
    rkey = e[:, 1] * (nv + 1) + e[:, 0]
    del e
    assert torch.unique(key).numel() == key.numel()
    assert torch.equal(torch.sort(key).values, torch.sort(rkey).values)
    assert bool(torch.isfinite(V).all())


@pytest.mark.parametrize("i", [0, 1])
def test_cranium_surface_matches_checker_and_envelope(sp, orc, cranium, i):
    """The reference's own Cranium masks (bit-packed golden): the device mesh equals the CPU checker's
    bit for bit and encloses the volume the reference recorded for its surface to within 2 %."""
    from test_oracle_mc import cranium_mask, mesh_volume_area
    sx, sy, sz = (float(v) for v in cranium["spacing"])
    pad = cranium_mask(cranium, i)
    V, F = sp.contour(pad, 127, (sx, sy, sz), 0, True, padding=(1, 1, 1))
    Vo, Fo = orc.marching_cubes(pad, 127, (sx, sy, sz), (-1, -1, -1), True)
    assert np.array_equal(F, Fo) and np.array_equal(V, Vo)
    vol, _ = mesh_volume_area(V, F.astype(np.int64))
    want = float(cranium[f"surface_{i}_volume_mm3"])
    assert abs(vol - want) / want < 0.02, (vol, want)
