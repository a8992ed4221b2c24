"""Watershed oracle pinning (CPU only).

scipy.ndimage.watershed_ift IS the reference's callee (invesalius/data/watershed_process.py:46,57)
and is present here, so it is the oracle. Two things are established on top of it:
  1. oracle/watershed.c::orc_ift_scipy restates SciPy's NI_WatershedIFT pointer for pointer
     (stack-like buckets, flat-index neighbourhood, the `if (p->next || p->prev)` unlink quirk) and
     reproduces SciPy BIT FOR BIT — i.e. SciPy's queue order is understood, not guessed;
  2. the order-independence model (orc_ws_model): wherever every chain of cost-optimal
     predecessors carries one label, that label IS SciPy's answer — in every case in which the
     unlink quirk does not corrupt SciPy's queues (then SciPy's costs are not minimax any more).
The GPU flood computes exactly this model (tests/test_gpu_watershed.py), so its parity with the
reference is exact on the order-independent voxels and a documented tie rule elsewhere.
"""
import numpy as np
import pytest
from scipy import ndimage
from scipy.ndimage import generate_binary_structure

from oracle import watershed as W


def random_case(rng, maxdim=9, signed=True):
    shape = tuple(int(rng.integers(2, maxdim)) for _ in range(3))
    img = rng.integers(0, int(rng.choice([3, 8, 40, 300, 65536])), shape).astype(np.uint16)
    mk = np.zeros(shape, np.int16)
    for lab in range(1, int(rng.integers(2, 5))):
        neg = signed and rng.random() < 0.15
        mk[tuple(int(rng.integers(s)) for s in shape)] = -lab if neg else lab
    st = generate_binary_structure(3, int(rng.choice([1, 2, 3])))
    return img, mk, st


def ct_case(shape=(48, 96, 96), seed=4):
    from invesalius3_b200 import phantom
    vol = phantom.ct(shape, seed=seed)
    rng = np.random.default_rng(seed)
    markers = np.zeros(vol.shape, np.uint8)
    inside, outside = np.argwhere(vol > 600), np.argwhere(vol < -900)
    for _ in range(4):
        z, y, x = inside[rng.integers(len(inside))]; markers[z, y, x] = 1
        z, y, x = outside[rng.integers(len(outside))]; markers[z, y, x] = 2
    return vol, markers


def test_restatement_reproduces_scipy_bit_for_bit():
    rng = np.random.default_rng(11)
    for _ in range(300):
        img, mk, st = random_case(rng)
        want = ndimage.watershed_ift(img, mk, st)
        got, _ = W.ift_scipy_restated(img, mk, st, quirk=True)
        assert np.array_equal(got, want)
    vol, markers = ct_case()
    pre = W.preprocess(vol, True, -18, 406)
    got, sole = W.ift_scipy_restated(pre, markers.astype(np.int16), generate_binary_structure(3, 1))
    assert np.array_equal(got, ndimage.watershed_ift(pre, markers.astype(np.int16), generate_binary_structure(3, 1)))
    assert sole > 0     # the quirk's precondition occurs all the time; it rarely changes a label


def test_order_independent_voxels_are_scipys_answer_unless_the_quirk_fires():
    rng = np.random.default_rng(12)
    quirk_cases = violations_in_quirk_cases = 0
    for _ in range(300):
        img, mk, st = random_case(rng)
        want = ndimage.watershed_ift(img, mk, st)
        clean, _ = W.ift_scipy_restated(img, mk, st, quirk=False)
        _, sets = W.order_independence_model(img, mk, st, 0)
        ok = sets != W.MULTI
        assert np.array_equal(sets[mk != 0], mk[mk != 0])
        if np.array_equal(clean, want):
            assert np.array_equal(want[ok], sets[ok].astype(np.int16))      # exact, no tolerance
        else:
            quirk_cases += 1
            violations_in_quirk_cases += int((want[ok] != sets[ok]).sum())
        # the algorithm as intended always agrees with the model
        assert np.array_equal(clean[ok], sets[ok].astype(np.int16))
    print(f"cases where SciPy's unlink quirk changes its result: {quirk_cases}/300 "
          f"({violations_in_quirk_cases} voxels off the minimax labelling)")
    assert quirk_cases < 40


def test_skimage_model_on_restated_flood():
    rng = np.random.default_rng(13)
    for _ in range(200):
        img, mk, st = random_case(rng, signed=False)
        want = W.watershed_skimage(img, mk, st)
        _, sets = W.order_independence_model(img, mk, st, 1)
        ok = sets != W.MULTI
        assert np.array_equal(want[ok], sets[ok].astype(np.int16))


@pytest.mark.parametrize("algorithm,mode", [("Watershed IFT", 0), ("Watershed", 1)])
def test_ct_phantom_is_almost_entirely_order_independent(algorithm, mode):
    """BASELINE configs[3] in small: with the Cranium window the labels are decided by cost
    almost everywhere, so the reference's answer does not depend on its queue order there."""
    vol, markers = ct_case()
    st = generate_binary_structure(3, 1)
    pre = W.preprocess(vol, True, -18, 406)
    if algorithm == "Watershed":
        pre = ndimage.morphological_gradient(pre, 3)
    want = W.do_watershed_array(vol, markers, st, algorithm, 3, True, -18, 406)
    _, sets = W.order_independence_model(pre, markers.astype(np.int16), st, mode)
    ok = sets != W.MULTI
    assert np.array_equal(want[ok], sets[ok].astype(want.dtype))
    assert ok.mean() > 0.99
