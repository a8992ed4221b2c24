"""CPU checker of apply_view_matrix_transform (oracle/transforms.c: transforms.rs:9-55 + interpolation.rs
restated; the reference has no test for it — parity unpinned) against closed-form cases."""
import numpy as np
import pytest


def _vol(shape=(12, 14, 16), seed=0):
    return np.random.default_rng(seed).integers(-1000, 3000, shape).astype(np.int16)


@pytest.mark.parametrize("interp", [0, 1, 2, 3])
def test_identity_matrix_returns_the_slab(orc, interp):
    """M = I: every sample falls on a voxel centre; all four interpolators reproduce it there (the
    Lanczos kernel is 1 at 0 and 0 at the other integers up to rounding), except on the last
    row / column / plane, which lies outside [0, d - 1) and takes cval."""
    vol = _vol()
    for orient, n, oshape in (("AXIAL", 4, (3, 14, 16)), ("CORONAL", 5, (12, 2, 16)), ("SAGITAL", 3, (12, 14, 4))):
        out = np.zeros(oshape, np.int16)
        orc.apply_view_matrix_transform(vol, (1.0, 0.5, 2.0), np.eye(4), n, orient, interp, -1024, out)   # spacings exact in binary
        z0, y0, x0 = (n if orient == "AXIAL" else 0), (n if orient == "CORONAL" else 0), (n if orient == "SAGITAL" else 0)
        want = vol[z0:z0 + oshape[0], y0:y0 + oshape[1], x0:x0 + oshape[2]].copy()
        zz, yy, xx = np.ogrid[z0:z0 + oshape[0], y0:y0 + oshape[1], x0:x0 + oshape[2]]
        inside = (zz < 11) & (yy < 13) & (xx < 15)
        assert np.array_equal(out[inside], want[inside]) if interp != 3 else np.abs(out[inside].astype(int) - want[inside]).max() <= 1
        assert (out[~inside] == -1024).all()


def test_integer_translation_and_dtypes(orc):
    """A translation by whole voxels: nearest and trilinear both pick the shifted voxel; uint8 and
    float64 volumes go through the same arithmetic."""
    vol = _vol()
    sp = (1.0, 0.5, 2.0)        # exact in binary: z * s / s is the integer again
    M = np.eye(4); M[0, 3] = 2 * sp[2]; M[1, 3] = -1 * sp[1]; M[2, 3] = 3 * sp[0]     # (z, y, x) order of the reference
    for a in (vol, (vol // 16 + 64).astype(np.uint8), vol.astype(np.float64) * 0.5):
        for interp in (0, 1):
            out = np.zeros((2, 14, 16), a.dtype)
            orc.apply_view_matrix_transform(a, sp, M, 1, "AXIAL", interp, 7, out)
            want = np.full(out.shape, 7, a.dtype)
            want[:, 1:14, 0:12] = a[3:5, 0:13, 3:15]
            assert np.array_equal(out, want), (a.dtype, interp)


def test_unrepresentable_value_is_reported(orc):
    vol = np.zeros((6, 6, 6), np.uint8)
    vol[2:4, 2:4, 2:4] = 255
    M = np.eye(4); M[2, 3] = 0.5            # half a voxel along x: the cubic overshoots 255 / undershoots 0
    with pytest.raises(ValueError):
        orc.apply_view_matrix_transform(vol, (1.0, 1.0, 1.0), M, 0, "AXIAL", 2, 0, np.zeros((6, 6, 6), np.uint8))
