"""VolumeSession (device-resident threshold -> region grow -> surface) gives exactly what the one-shot
numpy-in / numpy-out functions give, into the same host arrays."""
import numpy as np
import pytest
from scipy.ndimage import generate_binary_structure

pytestmark = pytest.mark.gpu


def test_session_equals_one_shot_calls(orc):
    from invesalius3_b200 import invesalius_rs as rs, phantom, slice_ops, surface_process
    from invesalius3_b200.session import VolumeSession
    vol = phantom.ct((40, 56, 72), seed=3)
    dz, dy, dx = vol.shape
    thr = (226, 3071)
    st = generate_binary_structure(3, 1)
    seed = phantom.first_seed_in_range(vol, dz // 2, *thr)
    mm1 = np.zeros((dz + 1, dy + 1, dx + 1), np.uint8); mm2 = mm1.copy()
    out1 = np.zeros(vol.shape, np.uint8); out2 = out1.copy()
    slice_ops.set_mask_threshold(vol, mm1, thr)
    rs.floodfill_threshold(vol, [seed], thr[0], thr[1], 254, st, out1)
    v1, f1 = surface_process.contour(out1, [127], (0.9, 0.9, 1.5), 0, True)
    with VolumeSession(vol) as s:
        s.set_mask_threshold(mm2, thr)
        s.floodfill_threshold([seed], thr[0], thr[1], 254, st, out2)
        v2, f2 = s.contour([127], (0.9, 0.9, 1.5), 0, True)
        # a second action in the same session: other seed, previous content consulted
        pre = np.zeros(vol.shape, np.uint8); pre[:, dy // 2, :] = 254          # a wall already filled
        want = pre.copy()
        orc.floodfill_threshold(vol, [seed], thr[0], thr[1], 254, st, want)
        got = pre.copy()
        s.floodfill_threshold([seed], thr[0], thr[1], 254, st, got, out_has_content=True)
        assert np.array_equal(got, want)
        vi, fi = s.contour([226, 3071], (1, 1, 1), 0, True, source="image")
        vw, fw = surface_process.contour(vol, [226, 3071], (1, 1, 1), 0, True)
        assert np.array_equal(vi, vw) and np.array_equal(fi, fw)
        with pytest.raises(IndexError):
            s.floodfill_threshold([(dx, 0, 0)], thr[0], thr[1], 254, st, out2.copy())
        with pytest.raises(OverflowError):
            s.floodfill_threshold([seed], thr[0], 40000, 254, st, out2.copy())
    assert np.array_equal(mm1, mm2) and np.array_equal(out1, out2)
    assert np.array_equal(v1, v2) and np.array_equal(f1, f2)
    want = np.zeros(vol.shape, np.uint8)
    orc.floodfill_threshold(vol, [seed], thr[0], thr[1], 254, st, want)
    assert np.array_equal(out2, want)
