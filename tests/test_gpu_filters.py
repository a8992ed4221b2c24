"""Pre-filters and mask algebra against SciPy / NumPy, the reference's own callees."""
import numpy as np
import pytest
from scipy import ndimage

pytestmark = pytest.mark.gpu


def _img(shape, seed):
    rng = np.random.default_rng(seed)
    return (ndimage.gaussian_filter(rng.normal(size=shape), 1.0) * 3000 + rng.normal(size=shape) * 200).astype(np.int16)


@pytest.mark.parametrize("shape", [(9, 10, 11), (20, 33, 47), (1, 30, 40), (40, 64, 64)])
def test_median_and_mean_equal_scipy(shape):
    from invesalius3_b200 import filters
    a = _img(shape, sum(shape))
    for value in (1.0, 1.4, 1.5, 2.0, 3.0):      # sizes 3, 3, 4, 5, 5 (capped)
        size = max(3, min(int(2 * value + 1), 5))
        assert np.array_equal(filters.median_blur_filter(a, value), ndimage.median_filter(a, size=size)), (shape, value)
    for value in (0.5, 1.0, 2.0, 3.0):           # sizes 2, 3, 5, 7
        want = ndimage.uniform_filter(a, size=int(2 * value + 1)).astype(a.dtype)
        assert np.array_equal(filters.mean_blur_filter(a, value), want), (shape, value)


def test_gaussian_family_equals_scipy():
    """filters.py:5-66 statement for statement against SciPy on the host."""
    from invesalius3_b200 import filters
    for shape in ((12, 13, 14), (20, 33, 47), (1, 30, 40)):
        matrix = _img(shape, 3 + sum(shape))
        for sigma in (0.5, 1.0, 1.7, 3.0):
            assert np.array_equal(filters.gaussian_blur_filter(matrix, sigma), ndimage.gaussian_filter(matrix, sigma=sigma))
            assert np.array_equal(filters.despeckle_filter(matrix, sigma), ndimage.gaussian_filter(matrix, sigma=sigma))
        for value in (1.0, 2.5):
            dtype = matrix.dtype
            min_val, max_val = matrix.min(), matrix.max()
            float_matrix = matrix.astype(float)
            blurred = ndimage.gaussian_filter(float_matrix, sigma=1.0)
            detail = float_matrix - blurred
            sharpened = float_matrix + value * 0.5 * detail
            want = np.clip(sharpened, min_val, max_val).astype(dtype)
            assert np.array_equal(filters.sharpening_filter(matrix, value), want), (shape, value)
        for value, normalize in ((1.0, True), (2.0, True), (1.0, False)):
            f = ndimage.gaussian_filter(matrix.astype(float), sigma=value)
            sx, sy, sz = ndimage.sobel(f, axis=0), ndimage.sobel(f, axis=1), ndimage.sobel(f, axis=2)
            magnitude = np.sqrt(sx**2 + sy**2 + sz**2)
            if normalize:
                min_val, max_val = float(matrix.min()), float(matrix.max())
                mag_min = magnitude.min()
                mag_range = magnitude.max() - mag_min
                if mag_range > 0:
                    magnitude = (magnitude - mag_min) / mag_range * (max_val - min_val) + min_val
            with np.errstate(invalid="ignore"):
                want = magnitude.astype(matrix.dtype)
            assert np.array_equal(filters.border_detection_filter(matrix, value, normalize), want), (shape, value, normalize)


def test_boolean_ops_and_convolve_non_zero():
    from invesalius3_b200 import filters
    rng = np.random.default_rng(2)
    shape = (17, 20, 33)
    vals = np.array([0, 1, 2, 253, 254, 255], np.uint8)
    m1, m2 = vals[rng.integers(0, 6, shape)], vals[rng.integers(0, 6, shape)]
    want = {filters.BOOLEAN_UNION: ((m1 > 2) + (m2 > 2)) * 255, filters.BOOLEAN_DIFF: ((m1 > 2) ^ ((m1 > 2) & (m2 > 2))) * 255,
            filters.BOOLEAN_AND: ((m1 > 2) & (m2 > 2)) * 255, filters.BOOLEAN_XOR: np.logical_xor((m1 > 2), (m2 > 2)) * 255}
    for op, w in want.items():
        mm = np.ones((shape[0] + 1, shape[1] + 1, shape[2] + 1), np.uint8)      # future_mask.matrix[:] = 1
        filters.boolean_op(op, m1, m2, mm[1:, 1:, 1:])
        assert np.array_equal(mm[1:, 1:, 1:], w.astype(np.uint8)), op
        assert (mm[0] == 1).all() and (mm[:, 0] == 1).all()
    # calc_mask_area's kernel (slice_.py:2306-2317), restated loop of transforms_py.rs:66-88
    sx, sy, sz = 0.9, 0.8, 1.5
    k = np.zeros((3, 3, 3))
    k[1, 1, 1] = 2 * sx * sy + 2 * sx * sz + 2 * sy * sz
    k[0, 1, 1] = k[2, 1, 1] = -(sx * sy); k[1, 0, 1] = k[1, 2, 1] = -(sx * sz); k[1, 1, 0] = k[1, 1, 2] = -(sy * sz)
    vol = (ndimage.gaussian_filter(rng.normal(size=shape), 1.5) > 0) * 1.0
    got = filters.convolve_non_zero(vol, k, 1)
    pad = np.pad(vol, 1, constant_values=1.0)
    want = np.zeros(shape)
    for kk in range(3):
        for j in range(3):
            for i in range(3):
                want = want + pad[kk:kk + shape[0], j:j + shape[1], i:i + shape[2]] * k[kk, j, i]   # same order as the reference's loops
    want[vol == 0] = 0.0
    assert np.array_equal(got, want)
