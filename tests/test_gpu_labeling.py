"""GPU connected components against SciPy itself (the reference's callee: mask.py:526-530)."""
import numpy as np
import pytest
from scipy import ndimage
from scipy.ndimage import generate_binary_structure

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("conn", [1, 2, 3])
def test_label_equals_scipy(conn):
    from invesalius3_b200 import labeling
    rng = np.random.default_rng(conn)
    for shape, sigma, thr in (((7, 9, 11), 0.0, 0.3), ((33, 65, 129), 1.2, 0.0), ((40, 64, 64), 2.0, 0.1), ((1, 50, 70), 1.0, 0.0)):
        f = rng.normal(size=shape)
        if sigma:
            f = ndimage.gaussian_filter(f, sigma)
            f /= np.abs(f).max()
        img = f > thr * (1 if sigma else 1)
        st = generate_binary_structure(3, conn)
        want, n = ndimage.label(img, st, output=np.uint32)
        got, m = labeling.label(img, st, output=np.uint32)
        assert m == n and got.dtype == np.uint32
        assert np.array_equal(got, want), (shape, conn)
    # 2-D input as the per-slice tools pass it (mask.py:539-552)
    img2 = ndimage.gaussian_filter(rng.normal(size=(90, 110)), 1.5) > 0
    st2 = generate_binary_structure(2, min(conn, 2))
    want, n = ndimage.label(img2, st2, output=np.uint32)
    got, m = labeling.label(img2, st2)
    assert m == n and np.array_equal(got, want)
    with pytest.raises(ValueError):          # SciPy: structuring element is not symmetric
        bad = np.zeros((3, 3, 3), bool); bad[1, 1, 1] = bad[0, 1, 1] = True
        labeling.label(np.ones((4, 4, 4), bool), bad)


def test_label_worst_cases():
    """One giant component (a serpentine that visits every row), a checkerboard (every voxel its own
    component under 6-connectivity, one component under 26), all background, all foreground."""
    from invesalius3_b200 import labeling
    snake = np.zeros((6, 40, 64), bool)
    snake[:, ::2, :] = True
    snake[:, 1::4, -1] = True; snake[:, 3::4, 0] = True
    snake[:-1:2, -2:, 5] = True; snake[1::2, :2, 9] = True
    snake[1:, 0, 3] |= True
    zz, yy, xx = np.indices((9, 10, 11))
    checker = (zz + yy + xx) % 2 == 0
    for img in (snake, checker, np.zeros((5, 6, 7), bool), np.ones((5, 6, 7), bool)):
        for conn in (1, 3):
            st = generate_binary_structure(3, conn)
            want, n = ndimage.label(img, st, output=np.uint32)
            got, m = labeling.label(img, st)
            assert m == n and np.array_equal(got, want), conn


def test_count_regions_largest_component_and_fill_holes_auto(orc):
    from invesalius3_b200 import invesalius_rs as rs, labeling
    rng = np.random.default_rng(7)
    mask = (ndimage.gaussian_filter(rng.normal(size=(30, 48, 64)), 1.0) > 0.02).astype(np.uint8) * 255
    lab, n = ndimage.label(mask != 0, output=np.uint32)
    # count_regions(labels as int16, n): out[p] = size of p's region (count_regions.rs:5-18)
    img = lab.astype(np.int16)
    counts = np.bincount(img.ravel(), minlength=n + 1)
    assert np.array_equal(rs.count_regions(img, n), counts[img].astype(np.uint32))
    with pytest.raises(ValueError):
        rs.count_regions(img, 3)          # values above number_regions index out of bounds in the reference
    # get_largest_connected_component (imagedata_utils.py:717-721)
    want = lab == np.argmax(np.bincount(lab.flat)[1:]) + 1
    assert np.array_equal(labeling.get_largest_connected_component(mask != 0), want)
    # Mask.fill_holes_auto, 3-D: SciPy labelling + the checker's fill against the device-only path
    for conn in (6, 26):
        for size in (5, 200):
            st = generate_binary_structure(3, {6: 1, 26: 3}[conn])
            want_m = mask.copy()
            l2, n2 = ndimage.label(~(want_m > 127), st, output=np.uint32)
            ret_w = orc.fill_holes_automatically(want_m, l2, n2, size)
            got_m = mask.copy()
            ret_g = labeling.fill_holes_auto(got_m, conn, size)
            assert ret_g == ret_w and np.array_equal(got_m, want_m), (conn, size)


def test_label_512_equals_scipy_on_a_slab():
    """A 96 x 512 x 512 slab of the bench phantom's threshold mask (SciPy needs a second for it)."""
    from invesalius3_b200 import labeling, phantom
    vol = phantom.ct((96, 512, 512), seed=2)
    img = (vol >= 226) & (vol <= 3071)
    st = generate_binary_structure(3, 1)
    want, n = ndimage.label(img, st, output=np.uint32)
    got, m = labeling.label(img, st)
    assert m == n and np.array_equal(got, want)
