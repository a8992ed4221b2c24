"""GPU parity: threshold and MaxIP/MinIP/MeanIP vs the oracle / NumPy, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1, 1), (7, 9, 11), (3, 16, 64), (33, 65, 129), (40, 96, 128)]


@pytest.fixture(scope="module")
def dev():
    import torch
    from invesalius3_b200 import device
    device.require_cuda()
    return device


def _rand_i16(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(-1100, 3200, shape).astype(np.int16)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("preserve", [False, True])
def test_threshold_matches_oracle(dev, orc, shape, preserve):
    import torch
    img = _rand_i16(shape, 1)
    rng = np.random.default_rng(2)
    old = rng.choice(np.array([0, 1, 2, 3, 127, 128, 252, 253, 254, 255], np.uint8), size=shape)
    for lo, hi in [(226, 3071), (-142, 2986), (500, 400), (-40000, 40000), (3071, 3071), (-32768, -1100)]:
        want = old.copy()
        orc.threshold(img, lo, hi, want, preserve)
        t_img = torch.from_numpy(img).cuda()
        t_out = torch.from_numpy(old.copy()).cuda()
        got = dev.threshold(t_img, lo, hi, out=t_out, preserve_markers=preserve).cpu().numpy()
        assert np.array_equal(got, want), (shape, preserve, lo, hi)


def test_threshold_unaligned_views(dev, orc):
    import torch
    img = _rand_i16((5, 33, 77), 5)
    base = torch.from_numpy(np.concatenate([[0], img.ravel()]).astype(np.int16)).cuda()
    view = base[1:].view(img.shape)  # 2-byte misaligned start
    assert view.data_ptr() % 16 != 0
    got = dev.threshold(view, 226, 3071).cpu().numpy()
    want = np.zeros(img.shape, np.uint8)
    orc.threshold(img, 226, 3071, want, False)
    assert np.array_equal(got, want)


def test_threshold_cranium_golden(dev, cranium):
    import torch
    img = cranium["matrix_crop"]
    t = torch.from_numpy(img).cuda()
    for i in (0, 1):
        lo, hi = (int(v) for v in cranium[f"thr_{i}"])
        want = np.unpackbits(cranium[f"mask_{i}_crop_bits"])[: img.size].reshape(img.shape) * np.uint8(255)
        assert np.array_equal(dev.threshold(t, lo, hi).cpu().numpy(), want)


def test_threshold_masklayout_and_host_shims(dev, orc):
    import torch
    from invesalius3_b200 import slice_ops
    rng = np.random.default_rng(7)
    vol = _rand_i16((9, 10, 13), 3)
    # device kernel on the padded layout
    mm = np.zeros((10, 11, 14), np.uint8)
    mm[3, 0, 0] = 1  # slice 2 already thresholded -> skipped when only_dirty
    mm[3, 1:, 1:] = 77
    mm[5, 2, 2] = 254
    want = mm.copy()
    orc.do_threshold_to_all_slices_numpy(vol, want, (226, 3071))
    t = torch.from_numpy(mm).cuda()
    dev.threshold_masklayout(torch.from_numpy(vol).cuda(), 226, 3071, t, True, True)
    assert np.array_equal(t.cpu().numpy(), want)
    # host shims with numpy in/out on the same layout (strided view packing)
    got = mm.copy()
    slice_ops.do_threshold_to_all_slices(vol, got, (226, 3071))
    assert np.array_equal(got, want)
    got = mm.copy(); want2 = mm.copy()
    orc.set_mask_threshold_numpy(vol, want2, (226, 3071))
    slice_ops.set_mask_threshold(vol, got, (226, 3071))
    assert np.array_equal(got, want2)
    old = rng.choice(np.array([0, 1, 2, 253, 254, 255], np.uint8), size=vol[0].shape)
    assert np.array_equal(slice_ops.do_threshold_to_a_slice(vol[0], old, (226, 3071)),
                          orc.do_threshold_to_a_slice_numpy(vol[0], old, (226, 3071)))
    assert np.array_equal(slice_ops.set_mask_threshold_slice(vol[0], (226, 3071)),
                          (255 * ((vol[0] >= 226) & (vol[0] <= 3071))).astype("uint8"))
    with pytest.raises(TypeError):
        slice_ops.set_mask_threshold_slice(vol[0].astype(np.float32), (226, 3071))


@pytest.mark.parametrize("shape", SHAPES + [(64, 128, 256), (130, 40, 72)])
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("kind", ["max", "min", "mean"])
def test_mip_matches_numpy(dev, shape, axis, kind):
    import torch
    img = _rand_i16(shape, 11)
    got = dev.mip(torch.from_numpy(img).cuda(), axis, kind).cpu().numpy()
    want = {"max": img.max, "min": img.min, "mean": img.mean}[kind](axis)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want), (shape, axis, kind)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_mip_uint8(dev, axis):
    import torch
    img = np.random.default_rng(4).integers(0, 256, (17, 33, 50)).astype(np.uint8)
    for kind in ("max", "min", "mean"):
        got = dev.mip(torch.from_numpy(img).cuda(), axis, kind).cpu().numpy()
        want = {"max": img.max, "min": img.min, "mean": img.mean}[kind](axis)
        assert np.array_equal(got, want)


def test_mip_large_properties(dev):
    """256^3 (L2-resident but multi-wave): idempotence + agreement across axes."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(5)
    t = torch.randint(-1024, 3072, (256, 256, 256), dtype=torch.int16, device="cuda", generator=g)
    for axis in (0, 1, 2):
        assert torch.equal(dev.mip(t, axis, "max"), t.amax(dim=axis))
        assert torch.equal(dev.mip(t, axis, "min"), t.amin(dim=axis))
        want = t.to(torch.int64).sum(dim=axis).to(torch.float64) / t.shape[axis]
        assert torch.equal(dev.mip(t, axis, "mean"), want)


def test_minmax(dev):
    import torch
    for shape in [(1, 1, 3), (7, 9, 11), (40, 96, 128)]:
        img = _rand_i16(shape, 9)
        mm = dev.minmax(torch.from_numpy(img).cuda()).cpu().numpy()
        assert mm.tolist() == [float(img.min()), float(img.max())]
    f = np.random.default_rng(1).normal(size=(5, 6, 7))
    mm = dev.minmax(torch.from_numpy(f).cuda()).cpu().numpy()
    assert mm.tolist() == [float(np.float32(f.min())), float(np.float32(f.max()))]


def test_set_mask_threshold_pipelined_path(dev, orc):
    """dz >= 32 takes the two-stream slab pipeline; pinned and pageable hosts, odd shapes."""
    import torch
    from invesalius3_b200 import slice_ops
    for shape, pinned in [((40, 33, 47), False), ((67, 16, 64), True)]:
        vol = _rand_i16(shape, 21)
        mm = np.zeros(tuple(s + 1 for s in shape), np.uint8)
        mm[0, :, :] = 9          # flags of the other orientations must survive
        mm[:, 0, 1:] = 7
        want = mm.copy()
        orc.set_mask_threshold_numpy(vol, want, (226, 3071))
        if pinned:
            hv = torch.from_numpy(vol).pin_memory().numpy()
            hm = torch.from_numpy(mm).pin_memory().numpy()
        else:
            hv, hm = vol, mm.copy()
        slice_ops.set_mask_threshold(hv, hm, (226, 3071))
        assert np.array_equal(hm, want), shape
