"""CPU-only checks of the host side: argument coercion / error behaviour of the
invesalius_rs mirror (raised before any device work), stride analysis of the packer,
Z-shard arithmetic, and that compute entry points fail loudly without a CUDA device."""
import numpy as np
import pytest


def test_box_pitches_recognises_memmap_style_views():
    from invesalius3_b200.device import _box_pitches
    m = np.zeros((11, 12, 13), np.uint8)
    assert _box_pitches(m) == (13, 12 * 13)
    assert _box_pitches(m[1:, 1:, 1:]) == (13, 12 * 13)          # mask.matrix[1:,1:,1:]
    assert _box_pitches(m[:, :, ::2]) is None                      # x not contiguous
    assert _box_pitches(m[::2]) == (13, 2 * 12 * 13)
    assert _box_pitches(m.transpose(2, 1, 0)) is None
    i = np.zeros((5, 6, 7), np.int16)
    assert _box_pitches(i[2:4]) == (14, 84)
    assert _box_pitches(i[0:1, 1:3]) is not None


def test_shim_type_and_range_errors_before_device_work():
    from invesalius3_b200 import invesalius_rs as rs
    st = np.ones((3, 3, 3), np.uint8)
    d16 = np.zeros((3, 4, 5), np.int16)
    out = np.zeros((3, 4, 5), np.uint8)
    with pytest.raises(TypeError):
        rs.floodfill_threshold(d16.astype(np.float32), [(0, 0, 0)], 0, 1, 1, st, out)
    with pytest.raises(TypeError):
        rs.floodfill_threshold(d16, [(0, 0, 0)], 0, 1, 1, st, out.astype(np.int16))
    with pytest.raises(OverflowError):
        rs.floodfill_threshold(d16, [(0, 0, 0)], 0, 70000, 1, st, out)      # t1 extracted as i16
    with pytest.raises(OverflowError):
        rs.floodfill_threshold(d16, [(0, 0, 0)], 0, 1, 300, st, out)        # fill: u8
    with pytest.raises(TypeError):
        rs.floodfill_threshold(np.zeros((3, 4, 5)), [(0, 0, 0)], 0, 1, 1, st, out)  # f64 data: fill becomes float
    with pytest.raises(TypeError):
        rs.floodfill_threshold_inplace(d16, [(0, 0, 0)], 0.5, 1, 1, st)
    with pytest.raises(TypeError):
        rs.fill_holes_automatically(out, np.zeros((3, 4, 5), np.int32), 1, 1)
    with pytest.raises(TypeError):
        rs.mida(d16, 0, 1, 1, np.zeros((4, 5), np.uint8))
    with pytest.raises(OverflowError):
        rs.mida(d16, 0, 40000, 1, np.zeros((4, 5), np.int16))
    with pytest.raises(TypeError):
        rs.lmip(d16, 0, 1, 2, np.zeros((4, 5), np.uint8))
    with pytest.raises(TypeError):
        rs.fast_countour_mip(d16, 1.0, 0, 1, 1, 0, np.zeros((4, 5), np.uint8))


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must raise, never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from invesalius3_b200 import invesalius_rs as rs, slice_ops, surface_process
    d16 = np.zeros((3, 4, 5), np.int16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rs.floodfill_threshold(d16, [(0, 0, 0)], 0, 1, 1, np.ones((3, 3, 3), np.uint8), np.zeros((3, 4, 5), np.uint8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        slice_ops.set_mask_threshold_slice(d16[0], (0, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        surface_process.contour(np.zeros((3, 4, 5), np.uint8), [127])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rs.mida(d16, 0, 1, 1, np.zeros((4, 5), np.int16))


def test_zshard_arithmetic():
    from invesalius3_b200.dist import ZShard
    DZ, W = 23, 4
    covered = []
    for r in range(W):
        s = ZShard(DZ, r, W)
        covered += list(range(s.z0, s.z1))
        assert s.ze0 == s.z0 - (1 if r > 0 else 0) and s.ze1 == s.z1 + (1 if r < W - 1 else 0)
        assert s.nz_ext == s.ze1 - s.ze0
        seeds = s.local_seeds([(1, 2, z) for z in range(DZ)])
        assert [z for (_, _, z) in seeds] == list(range(s.nz_ext))
    assert covered == list(range(DZ))
    with pytest.raises(IndexError):
        ZShard(DZ, 0, W).local_seeds([(0, 0, DZ)])


def test_phantom_is_shard_consistent():
    from invesalius3_b200 import phantom
    a = phantom.ct((20, 24, 28), seed=5)
    assert a.dtype == np.int16 and a.min() >= -1024 and a.max() <= 3071
    assert np.array_equal(phantom.ct((20, 24, 28), seed=5, zrange=(7, 13)), a[7:13])
    assert not np.array_equal(phantom.ct((20, 24, 28), seed=6), a)


def test_vtp_round_trip_and_layout(tmp_path):
    """write_vtp emits what vtkXMLPolyDataWriter would for a triangle mesh (Points + Polys with
    connectivity / offsets, inline base64 with UInt32 byte counts); read_vtp inverts it."""
    import base64
    import xml.etree.ElementTree as ET
    from invesalius3_b200 import surface_process as sp
    rng = np.random.default_rng(0)
    v = rng.normal(size=(11, 3)).astype(np.float32)
    for dt in (np.int32, np.int64):
        f = rng.integers(0, 11, (7, 3)).astype(dt)
        fn = str(tmp_path / f"m_{np.dtype(dt).name}.vtp")
        sp.write_vtp(fn, v, f)
        v2, f2 = sp.read_vtp(fn)
        assert np.array_equal(v, v2) and np.array_equal(f, f2) and f2.dtype == dt
        root = ET.parse(fn).getroot()
        assert root.tag == "VTKFile" and root.get("type") == "PolyData" and root.get("byte_order") == "LittleEndian"
        piece = root.find("PolyData").find("Piece")
        assert piece.get("NumberOfPoints") == "11" and piece.get("NumberOfPolys") == "7"
        offs = [d for d in piece.find("Polys").findall("DataArray") if d.get("Name") == "offsets"][0]
        raw = base64.b64decode(offs.text.strip())
        assert int(np.frombuffer(raw[:4], np.uint32)[0]) == 7 * np.dtype(dt).itemsize
        assert np.array_equal(np.frombuffer(raw[4:], dt), np.arange(1, 8) * 3)
    sp.write_vtp(str(tmp_path / "empty.vtp"), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int64))
    v0, f0 = sp.read_vtp(str(tmp_path / "empty.vtp"))
    assert v0.shape == (0, 3) and f0.shape == (0, 3)


def test_non_hot_exports_are_forwarded_to_the_crate(monkeypatch):
    """invesalius_rs/__init__.py:273-300: names outside the hot path resolve to the reference crate
    when it is installed, and fail with a message naming the reason when it is not."""
    import sys
    import types
    from invesalius3_b200 import invesalius_rs as rs
    monkeypatch.setattr(rs, "_crate", None)
    monkeypatch.delitem(sys.modules, "invesalius_rs", raising=False)
    with pytest.raises(AttributeError, match="crate is not installed"):
        rs.mask_cut
    with pytest.raises(AttributeError):
        rs.no_such_symbol
    fake = types.ModuleType("invesalius_rs")
    fake.__file__ = "/somewhere/invesalius_rs/__init__.py"
    fake.mask_cut = lambda image, n: ("crate", n)
    fake.brush_mask_rs = type("B", (), {})
    monkeypatch.setitem(sys.modules, "invesalius_rs", fake)
    assert rs.mask_cut(None, 4) == ("crate", 4) and rs.brush_mask_rs is fake.brush_mask_rs
    for name in ("floodfill", "floodfill_threshold", "floodfill_threshold_inplace", "fill_holes_automatically", "mida",
                 "lmip", "fast_countour_mip"):
        assert callable(getattr(rs, name)) and getattr(rs, name).__module__ == rs.__name__   # hot path: ours
    assert set(rs.FORWARDED) >= {"mask_cut", "polygon2mask_rs", "brush_mask_rs"}
    for ours in ("apply_view_matrix_transform", "count_regions", "convolve_non_zero", "ca_smoothing", "Mesh"):   # SURVEY 8f-1 .. 8f-4
        assert ours not in rs.FORWARDED and callable(getattr(rs, ours))


def test_new_shims_reject_bad_arguments_before_device_work():
    """The 8f shims validate dtypes / ranges like the reference's PyO3 layer before touching the device."""
    from invesalius3_b200 import filters, invesalius_rs as rs, labeling, mesh_ops
    vol = np.zeros((4, 5, 6), np.int16)
    with pytest.raises(TypeError):        # volume / out dtype mismatch (transforms_py.rs:147)
        rs.apply_view_matrix_transform(vol, (1, 1, 1), np.eye(4), 0, "AXIAL", 1, 0, np.zeros((1, 5, 6), np.uint8))
    with pytest.raises(OverflowError):    # cval.extract::<i16>()
        rs.apply_view_matrix_transform(vol, (1, 1, 1), np.eye(4), 0, "AXIAL", 1, 40000, np.zeros((1, 5, 6), np.int16))
    with pytest.raises(TypeError):
        rs.apply_view_matrix_transform(vol, (1, 1), np.eye(4), 0, "AXIAL", 1, 0, np.zeros((1, 5, 6), np.int16))
    with pytest.raises(TypeError):
        filters.median_blur_filter(vol.astype(np.float32), 1.0)
    with pytest.raises(TypeError):
        filters.boolean_op(filters.BOOLEAN_AND, vol, vol, vol)
    with pytest.raises(KeyError):
        filters.boolean_op(9, vol.astype(np.uint8), vol.astype(np.uint8), vol.astype(np.uint8))
    with pytest.raises(OverflowError):
        filters.convolve_non_zero(np.zeros((3, 3, 3)), np.ones((3, 3, 3)), 70000)
    with pytest.raises(NotImplementedError):
        labeling.label(np.zeros((2, 2, 2, 2), bool))
    with pytest.raises(RuntimeError):
        labeling.label(np.zeros((4, 4), bool), np.ones((3, 3, 3), bool))
    with pytest.raises(TypeError):
        mesh_ops.context_aware_smoothing(np.zeros((3, 3), np.float64), np.zeros((1, 4), np.int64), np.zeros((1, 3), np.float32), 0.7, 3, 0.1, 1)
    with pytest.raises(TypeError):
        mesh_ops.context_aware_smoothing(np.zeros((3, 3), np.float32), np.zeros((1, 4), np.float32), np.zeros((1, 3), np.float32), 0.7, 3, 0.1, 1)
    with pytest.raises(ValueError):
        mesh_ops.Mesh()
    m = mesh_ops.Mesh(vertices=np.zeros((3, 3), np.float32), faces=np.array([[3, 0, 1, 2]]), normals=np.zeros((1, 3), np.float32))
    assert mesh_ops.Mesh(other=m).faces is not m.faces and rs.Mesh is mesh_ops.Mesh
