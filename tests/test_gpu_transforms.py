"""apply_view_matrix_transform on the device against the CPU restatement (oracle/transforms.c)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rotation(shape, sp, angles=(0.31, -0.12, 0.2)):
    """T1 * R * T0 about the volume centre in (z, y, x) millimetres, like slice_.py:2040-2045."""
    cz, cy, cx = shape[0] * sp[2] / 2, shape[1] * sp[1] / 2, shape[2] * sp[0] / 2
    a, b, c = angles
    Rz = np.array([[1, 0, 0, 0], [0, np.cos(a), -np.sin(a), 0], [0, np.sin(a), np.cos(a), 0], [0, 0, 0, 1.0]])
    Ry = np.array([[np.cos(b), 0, np.sin(b), 0], [0, 1, 0, 0], [-np.sin(b), 0, np.cos(b), 0], [0, 0, 0, 1.0]])
    Rx = np.array([[np.cos(c), -np.sin(c), 0, 0], [np.sin(c), np.cos(c), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    T0 = np.eye(4); T0[:3, 3] = (-cz, -cy, -cx)
    T1 = np.eye(4); T1[:3, 3] = (cz, cy, cx)
    return T1 @ Rz @ Ry @ Rx @ T0


@pytest.mark.parametrize("interp", [0, 1, 2, 3])
def test_rotated_slabs_match_checker(orc, interp):
    from invesalius3_b200 import invesalius_rs as rs
    from scipy import ndimage
    rng = np.random.default_rng(3)
    shape = (30, 41, 52)
    vol = (ndimage.gaussian_filter(rng.normal(size=shape), 1.0) * 2500).astype(np.int16)
    sp = (0.9570312, 0.9570312, 1.5)
    M = _rotation(shape, sp)
    cases = [(vol, "AXIAL", 11, (4, 41, 52), int(vol.min())), (vol, "CORONAL", 7, (30, 3, 52), -1024),
             (vol, "SAGITAL", 40, (30, 41, 2), 0), (vol, "VOLUME", 0, shape, int(vol.min())),
             ((vol // 32 + 100).clip(0, 255).astype(np.uint8), "AXIAL", 0, (30, 41, 52), 3),
             (vol.astype(np.float64) * 0.37, "AXIAL", 5, (6, 41, 52), -500.25)]
    for a, orient, n, oshape, cval in cases:
        want = np.zeros(oshape, a.dtype); got = want.copy()
        orc.apply_view_matrix_transform(a, sp, M, n, orient, interp, cval, want)
        rs.apply_view_matrix_transform(a, sp, M, n, orient, interp, cval, got)
        if interp < 3:
            assert np.array_equal(got, want), (a.dtype, orient, interp)
        else:       # sin(): libm and the device may differ in the last bit
            if a.dtype == np.float64:
                assert np.allclose(got, want, rtol=1e-12, atol=1e-9)
            else:
                diff = np.abs(got.astype(np.int64) - want.astype(np.int64))
                assert diff.max() <= 1 and (diff != 0).mean() < 1e-3, (orient, diff.max(), (diff != 0).mean())
        assert (want != cval).mean() > 0.3      # the rotated slab really samples the volume


def test_errors_like_the_reference(orc):
    from invesalius3_b200 import invesalius_rs as rs
    vol = np.zeros((6, 6, 6), np.uint8)
    vol[2:4, 2:4, 2:4] = 255
    M = np.eye(4); M[2, 3] = 0.5
    with pytest.raises(ValueError):           # NumCast::from(..).unwrap() panics in the reference
        rs.apply_view_matrix_transform(vol, (1.0, 1.0, 1.0), M, 0, "AXIAL", 2, 0, np.zeros((6, 6, 6), np.uint8))
    with pytest.raises(TypeError):
        rs.apply_view_matrix_transform(vol, (1.0, 1.0, 1.0), M, 0, "AXIAL", 1, 0, np.zeros((6, 6, 6), np.int16))
    with pytest.raises(OverflowError):
        rs.apply_view_matrix_transform(vol, (1.0, 1.0, 1.0), M, 0, "AXIAL", 1, 300, np.zeros((6, 6, 6), np.uint8))
    with pytest.raises(TypeError):
        rs.apply_view_matrix_transform(vol.astype(np.float32), (1.0, 1.0, 1.0), M, 0, "AXIAL", 1, 0, np.zeros((6, 6, 6), np.float32))
