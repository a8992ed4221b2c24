"""Watershed on the GPU.
  LUT / shift / morphological gradient: bit-exact against NumPy / SciPy (SciPy is the
  reference's own callee).
  Flood: SciPy's watershed_ift (true callee) and the restated skimage heap flood are
  sequential, queue-ordered algorithms. The GPU computes the exact minimax cost field, the SET
  of labels that can reach every voxel along cost-optimal edges, and a deterministic labelling.
  Asserted, with no tolerance:
    * the GPU's order-dependence mask equals the CPU model's (oracle/watershed.c::orc_ws_model);
    * on every order-INDEPENDENT voxel the GPU label equals the model's — and equals the
      reference's own output (SciPy / restated skimage) whenever SciPy's list-unlink quirk did
      not corrupt its queues (tests/test_oracle_watershed.py explains and pins that quirk).
  On order-dependent voxels (genuine queue-order ties) the agreement is measured, printed and
  bounded from below at the measured value minus one point.
"""
import multiprocessing
import os
import tempfile

import numpy as np
import pytest
from scipy import ndimage
from scipy.ndimage import generate_binary_structure

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wp():
    from invesalius3_b200 import device, watershed_process
    device.require_cuda()
    return watershed_process


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_lut_shift_bit_exact(wp):
    from oracle import watershed as W
    rng = np.random.default_rng(0)
    img = rng.integers(-1100, 3200, (9, 33, 70)).astype(np.int16)
    for ww, wl in [(406, -18), (300, 300), (2, 0), (4000, 1000), (255, 127)]:
        want = W.get_LUT_value(img, ww, wl).astype("uint16")
        got = wp.lut_u16(_t(img), ww, wl).cpu().numpy().view(np.uint16)
        assert np.array_equal(got, want), (ww, wl)
    want = (img - img.min()).astype("uint16")
    got = wp.shift_u16(_t(img)).cpu().numpy().view(np.uint16)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("size", [3, (3, 3, 3), (1, 3, 3), 5, (3, 5, 1), 2, (2, 3, 4)])
def test_morphological_gradient_matches_scipy(wp, size):
    rng = np.random.default_rng(1)
    for shape in [(7, 9, 11), (3, 4, 5), (20, 33, 65), (1, 5, 6)]:
        pre = rng.integers(0, 4000, shape).astype(np.uint16)
        want = ndimage.morphological_gradient(pre, size)
        got = wp.morphological_gradient_u16(_t(pre.view(np.int16)), size).cpu().numpy().view(np.uint16)
        assert want.dtype == np.uint16 and np.array_equal(got, want), (shape, size)


def _tie_free_image(shape, seed):
    """uint16 image whose pairwise differences are all distinct enough that no two paths
    from different markers tie: a random permutation scaled by large distinct steps."""
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    vals = rng.permutation(n).astype(np.int64)
    return (vals * (60000 // max(n, 1)) if n <= 60000 else vals % 60000).astype(np.uint16).reshape(shape)


def _markers(shape, seed, k=4):
    rng = np.random.default_rng(seed)
    m = np.zeros(shape, np.int16)
    for lab in range(1, k + 1):
        z, y, x = (int(rng.integers(s)) for s in shape)
        m[z, y, x] = lab
    return m


@pytest.mark.parametrize("conn", [1, 3])
def test_ift_matches_scipy_when_tie_free(wp, conn):
    """Random full-range uint16 volumes: edge weights |dI| rarely collide, so SciPy's answer
    is (almost) free of queue-order ties and must be reproduced — including its flat-array
    neighbourhood that wraps rows and planes into each other at the volume faces."""
    st = generate_binary_structure(3, conn)
    rng = np.random.default_rng(conn)
    agree_total, n_total = 0, 0
    for seed, shape in enumerate([(6, 7, 8), (10, 12, 14), (5, 30, 31), (20, 33, 47), (3, 70, 18)]):
        img = rng.integers(0, 65536, shape).astype(np.uint16)
        mk = _markers(shape, seed + 10)
        want = ndimage.watershed_ift(img, mk, st)
        got = wp.flood(_t(img.view(np.int16)), _t(mk), st, "Watershed IFT").cpu().numpy()
        agree_total += int((got == want).sum()); n_total += want.size
        assert np.array_equal(got[mk != 0], mk[mk != 0])
        assert (got != 0).all()
    assert agree_total / n_total >= 0.995, agree_total / n_total
    # narrow value range: many equal edge weights, i.e. genuine queue-order ties
    img = rng.integers(0, 300, (8, 20, 20)).astype(np.uint16)
    mk = _markers(img.shape, 99)
    want = ndimage.watershed_ift(img, mk, st)
    got = wp.flood(_t(img.view(np.int16)), _t(mk), st, "Watershed IFT").cpu().numpy()
    print(f"narrow-range (tie-heavy) IFT agreement, conn {conn}: {(got == want).mean():.4f}")
    assert (got == want).mean() >= FLOOR_NARROW[conn], (got == want).mean()


# measured agreement on inputs with genuine queue-order ties, minus one point (a regression
# below these values fails; the exact part of the parity is asserted separately, below)
FLOOR_NARROW = {1: 0.60, 3: 0.60}
FLOOR_CT = {"Watershed": 0.98, "Watershed IFT": 0.98}


def _cases(rng, n, signed=False):
    from test_oracle_watershed import random_case
    return [random_case(rng, maxdim=14, signed=signed) for _ in range(n)]


@pytest.mark.parametrize("algorithm,mode", [("Watershed IFT", 0), ("Watershed", 1)])
def test_exact_on_order_independent_voxels_random(wp, algorithm, mode):
    """Random volumes (3 .. 65536 grey levels, 1-4 markers, 6/18/26 neighbourhoods): the GPU's
    order-dependence mask equals the model's, its labels equal the model's on every
    order-independent voxel, and those equal the reference's output unless SciPy's unlink quirk
    corrupted its queues for that input."""
    from oracle import watershed as W
    rng = np.random.default_rng(100 + mode)
    checked = amb_total = n_total = quirk_cases = 0
    for img, mk, st in _cases(rng, 60):
        got, amb = wp.flood(_t(img.view(np.int16)), _t(mk), st, algorithm, return_ambiguous=True)
        got, amb = got.cpu().numpy(), amb.cpu().numpy().astype(bool)
        _, sets = W.order_independence_model(img, mk, st, mode)
        assert np.array_equal(amb, sets == W.MULTI), "order-dependence masks differ"
        ok = ~amb
        assert np.array_equal(got[ok], sets[ok].astype(np.int16)), "label differs on an order-independent voxel"
        assert (got != 0).all() and np.array_equal(got[mk != 0], mk[mk != 0])
        if mode == 0:
            want = ndimage.watershed_ift(img, mk, st)
            clean, _ = W.ift_scipy_restated(img, mk, st, quirk=False)
            if np.array_equal(clean, want):
                assert np.array_equal(got[ok], want[ok])
            else:
                quirk_cases += 1
        else:
            assert np.array_equal(got[ok], W.watershed_skimage(img, mk, st)[ok])
        checked += int(ok.sum()); amb_total += int(amb.sum()); n_total += img.size
    print(f"{algorithm}: {checked} order-independent voxels exact, {amb_total}/{n_total} order-dependent, "
          f"{quirk_cases} inputs where SciPy's unlink quirk fired")


@pytest.mark.parametrize("shape", [(48, 96, 96), (128, 128, 128)])
def test_exact_on_order_independent_voxels_ct(wp, shape):
    """BASELINE configs[3] in small (CT phantom, Cranium window ww 406 / wl -18, 6-connected):
    exact on every order-independent voxel against the reference's callee itself."""
    from oracle import watershed as W
    from test_oracle_watershed import ct_case
    vol, markers = ct_case(shape)
    st = generate_binary_structure(3, 1)
    for algorithm, mode in (("Watershed", 1), ("Watershed IFT", 0)):
        want = W.do_watershed_array(vol, markers, st, algorithm, 3, True, -18, 406)
        got, amb = wp.watershed_device(_t(vol), _t(markers), st, algorithm, 3, True, -18, 406, return_ambiguous=True)
        got, amb = got.cpu().numpy(), amb.cpu().numpy().astype(bool)
        pre = W.preprocess(vol, True, -18, 406)
        if algorithm == "Watershed":
            pre = ndimage.morphological_gradient(pre, 3)
        _, sets = W.order_independence_model(pre, markers.astype(np.int16), st, mode)
        assert np.array_equal(amb, sets == W.MULTI)
        assert np.array_equal(got[~amb], want[~amb]), f"{algorithm}: differs from the reference on an order-independent voxel"
        frac = float((got == want).mean())
        print(f"{algorithm} {shape}: order-independent {1 - amb.mean():.5f} (exact), overall agreement {frac:.5f}")
        assert frac >= FLOOR_CT[algorithm]


def test_ift_simple_known_answers(wp):
    """SURVEY appendix B behaviours that do not depend on queue order."""
    st = np.zeros((3, 3, 3), np.uint8); st[1, 1, :] = 1   # SciPy insists on 3 x 3 x 3 structures
    img = np.array([0, 0, 0, 0, 5, 0, 0, 0, 0], np.uint16).reshape(1, 1, 9)
    mk = np.zeros((1, 1, 9), np.int16); mk[0, 0, 0] = 1; mk[0, 0, 8] = 2
    want = ndimage.watershed_ift(img, mk, st)
    got = wp.flood(_t(img.view(np.int16)), _t(mk), st, "Watershed IFT").cpu().numpy()
    # the ridge voxel itself is a tie (cost 5 from both sides): compare the rest
    keep = np.ones(9, bool); keep[4] = False
    assert np.array_equal(got.ravel()[keep], want.ravel()[keep])


def test_value_flood_matches_restated_skimage_when_tie_free(wp):
    from oracle import watershed as W
    st = generate_binary_structure(3, 1)
    agree_total, n_total = 0, 0
    for seed, shape in enumerate([(6, 7, 8), (10, 12, 14), (4, 25, 26)]):
        img = _tie_free_image(shape, seed + 3)
        mk = _markers(shape, seed + 20)
        want = W.watershed_skimage(img, mk, st)
        got = wp.flood(_t(img.view(np.int16)), _t(mk), st, "Watershed").cpu().numpy()
        agree_total += int((got == want).sum()); n_total += want.size
        assert (got != 0).all()
    assert agree_total / n_total >= 0.98, agree_total / n_total


def test_do_watershed_reference_smoke(wp):
    """tests/test_segmentation_tools.py:170-213, same arguments, same assertions — and equal
    to the CPU checker on this input."""
    from oracle import watershed as W
    image = np.zeros((5, 5, 5), dtype=np.int16)
    image[1:4, 1:4, 1:4] = 100
    markers = np.zeros((5, 5, 5), dtype=np.int16)
    markers[2, 2, 2] = 1
    markers[0, 0, 0] = 2
    bstruct = generate_binary_structure(3, 1)
    for algorithm in ("Watershed", "Watershed IFT"):
        q = multiprocessing.Queue()
        with tempfile.TemporaryDirectory() as temp_dir:
            tfile = os.path.join(temp_dir, "watershed_mask.tmp")
            tmp_mask = np.memmap(tfile, shape=(5, 5, 5), dtype="uint8", mode="w+")
            wp.do_watershed(image=image, markers=markers, tfile=tfile, shape=(5, 5, 5), bstruct=bstruct,
                            algorithm=algorithm, mg_size=(3, 3, 3), use_ww_wl=False, wl=0, ww=0, q=q)
            result = tmp_mask.copy()
            tmp_mask._mmap.close()
            del tmp_mask
        assert np.any(result > 0), "Watershed should produce segmentation"
        assert q.get(timeout=2) == 1
        assert set(np.unique(result)) <= {1, 2}
        assert result[2, 2, 2] == 1 and result[0, 0, 0] == 2
        want = W.do_watershed_array(image, markers, bstruct, algorithm, (3, 3, 3), False, 0, 0)
        # the cube interior / exterior are decided by cost, not by queue order
        assert result[1:4, 1:4, 1:4].min() == 1 or algorithm == "Watershed"
        agree = (result == want.astype(np.uint8)).mean()
        assert agree >= 0.75, (algorithm, agree)   # 125 voxels, most of them plateau ties


def test_ct_phantom_agreement_fraction(wp):
    """CT-like input with LUT plateaus (configs[3] in small): report how much of the volume
    gets the reference's label. Plateau ties are resolved differently by construction."""
    import torch
    from oracle import watershed as W
    from invesalius3_b200 import phantom
    vol = phantom.ct((48, 96, 96), seed=4)
    rng = np.random.default_rng(4)
    markers = np.zeros(vol.shape, np.uint8)
    inside = np.argwhere(vol > 600)
    outside = np.argwhere(vol < -900)
    for k in range(4):
        z, y, x = inside[rng.integers(len(inside))]; markers[z, y, x] = 1
        z, y, x = outside[rng.integers(len(outside))]; markers[z, y, x] = 2
    st = generate_binary_structure(3, 1)
    for algorithm in ("Watershed", "Watershed IFT"):
        want = W.do_watershed_array(vol, markers, st, algorithm, 3, True, -18, 406)
        got = wp.watershed_device(_t(vol), _t(markers), st, algorithm, 3, True, -18, 406).cpu().numpy()
        frac = float((got == want).mean())
        print(f"agreement[{algorithm}] = {frac:.4f}")
        assert set(np.unique(got)) <= {1, 2}
        assert frac >= FLOOR_CT[algorithm], (algorithm, frac)


# ---- the persistent 6-connected engine against the NumPy model of the same definition --------------
@pytest.mark.parametrize("shape", [(7, 9, 11), (20, 33, 48), (18, 32, 64), (1, 40, 50)])
@pytest.mark.parametrize("mode", [0, 1])
def test_flood_equals_numpy_model(wp, shape, mode):
    """Labels AND the order-dependence mask equal the whole-array NumPy restatement (tests/ws_model.py)
    bit for bit: partial tiles (generic path), full aligned tiles (register chains), a 2-D slice."""
    import torch
    import ws_model
    rng = np.random.default_rng(sum(shape) + mode)
    img = ndimage.gaussian_filter(rng.normal(size=shape), 1.2)
    img = ((img - img.min()) / (np.ptp(img) + 1e-9) * 300).astype(np.uint16)     # plateaus and ties on purpose
    mk = np.zeros(shape, np.int16)
    for lab in (1, 2, -3, 2):
        z, y, x = (int(rng.integers(0, s)) for s in shape)
        mk[z, max(0, y - 1): y + 2, max(0, x - 1): x + 2] = lab
    st = generate_binary_structure(3, 1)
    lab, amb = wp.flood(_t(img.view(np.int16)), _t(mk), st, "Watershed" if mode == 1 else "Watershed IFT", True)
    want_lab, want_amb = ws_model.flood(img, mk, mode)
    assert np.array_equal(lab.cpu().numpy(), want_lab)
    assert np.array_equal(amb.cpu().numpy(), want_amb)
    # labels without the set propagation (what do_watershed runs) are the same labels
    lab2 = wp.flood(_t(img.view(np.int16)), _t(mk), st, "Watershed" if mode == 1 else "Watershed IFT")
    assert np.array_equal(lab2.cpu().numpy(), want_lab)


def test_engine_matches_generic_kernels(wp, monkeypatch):
    """The same structuring element through the generic round kernels (B2V_WS_GENERIC) gives the
    same labels and mask."""
    import torch
    from invesalius3_b200 import phantom
    vol = phantom.ct((40, 48, 64), seed=4)
    mk = np.zeros(vol.shape, np.uint8)
    mk[20, 24, 30:34] = 1; mk[2, 2, 2:6] = 2; mk[30, 40, 50:54] = 1; mk[38, 5, 60:63] = 2
    st = generate_binary_structure(3, 1)
    res = {}
    for tag in ("fast", "generic"):
        if tag == "generic":
            monkeypatch.setenv("B2V_WS_GENERIC", "1")
        for alg in ("Watershed", "Watershed IFT"):
            r = wp.watershed_device(_t(vol), _t(mk), st, alg, 3, True, -18, 406, return_ambiguous=True)
            res[(tag, alg)] = (r[0].cpu().numpy(), r[1].cpu().numpy())
    for alg in ("Watershed", "Watershed IFT"):
        assert np.array_equal(res[("fast", alg)][0], res[("generic", alg)][0])
        assert np.array_equal(res[("fast", alg)][1], res[("generic", alg)][1])


def test_many_labels_take_the_64_bit_keys(wp):
    """More than 256 distinct marker labels do not fit the ranked 32-bit key (hops << 8 | rank): the
    one-shot call then runs phase 2 on 64-bit keys. Same model, same answer; negative labels included."""
    import ws_model
    rng = np.random.default_rng(5)
    shape = (20, 33, 48)
    img = (ndimage.gaussian_filter(rng.normal(size=shape), 1.0) * 200 + 300).clip(0, 600).astype(np.uint16)
    for nlab in (200, 300):
        mk = np.zeros(shape, np.int16)
        pos = rng.choice(img.size, nlab, replace=False)
        labs = np.concatenate([np.arange(1, nlab // 2 + 1), -np.arange(1, nlab - nlab // 2 + 1)]).astype(np.int16)
        mk.ravel()[pos] = labs
        st = generate_binary_structure(3, 1)
        for mode, alg in ((0, "Watershed IFT"), (1, "Watershed")):
            lab, amb = wp.flood(_t(img.view(np.int16)), _t(mk), st, alg, True)
            want_lab, want_amb = ws_model.flood(img, mk, mode)
            assert np.array_equal(lab.cpu().numpy(), want_lab), (nlab, alg)
            assert np.array_equal(amb.cpu().numpy(), want_amb), (nlab, alg)
