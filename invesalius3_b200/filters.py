"""Pre-filters and mask algebra with the reference's signatures, numpy in / numpy out:

  median_blur_filter(matrix, value)   invesalius/data/filters.py:9-12   ndimage.median_filter, size 3 or 5
  mean_blur_filter(matrix, value)     filters.py:15-18                  ndimage.uniform_filter(...).astype(dtype)
  boolean_op(op, m1, m2, out)         Slice.do_boolean_op, slice_.py:1906-1916 (mask bodies)
  convolve_non_zero(volume, kernel, cval)   invesalius_rs.convolve_non_zero (calc_mask_area, slice_.py:2299-2322)

Bit-exact against SciPy / NumPy. The Gaussian-based filters of filters.py (gaussian_blur, sharpening,
despeckle, border detection) are not built: SciPy's float kernels would have to be reproduced to the
bit; they raise NotImplementedError here rather than fall back to the CPU.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from . import device as dev
from .device import _p, _stream

BOOLEAN_UNION, BOOLEAN_DIFF, BOOLEAN_AND, BOOLEAN_XOR = 1, 2, 3, 4     # invesalius/constants.py:818-821


def _i16_volume(matrix):
    a = np.asarray(matrix)
    if a.dtype != np.int16 or a.ndim != 3:
        raise TypeError("filter: int16 3-D matrix expected")
    return a


def median_blur_filter(matrix: np.ndarray, value: float) -> np.ndarray:
    a = _i16_volume(matrix)
    size = max(3, min(int(2 * value + 1), 5))
    if size == 4:
        raise NotImplementedError("median_filter: even sizes are not built (the reference reaches 3, 4 or 5)")
    t = dev.to_device(a)
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        _lib.call("b2v_median_filter_i16", _p(t), *a.shape, size, _p(out), _stream())
    res = np.empty(a.shape, np.int16)
    dev.to_host(out, res)
    return res


def mean_blur_filter(matrix: np.ndarray, value: float) -> np.ndarray:
    a = _i16_volume(matrix)
    size = int(2 * value + 1)
    if size < 1:
        raise RuntimeError("incorrect filter size")      # SciPy's message
    t = dev.to_device(a)
    out, tmp = torch.empty_like(t), torch.empty_like(t)
    with torch.cuda.device(t.device):
        _lib.call("b2v_uniform_filter_i16", _p(t), *a.shape, size, _p(out), _p(tmp), _stream())
    res = np.empty(a.shape, np.int16)
    dev.to_host(out, res)
    return res


def _not_built(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name}: Gaussian-based filters are not built on the device (no CPU fallback)")
    f.__name__ = name
    return f


gaussian_blur_filter = _not_built("gaussian_blur_filter")
sharpening_filter = _not_built("sharpening_filter")
despeckle_filter = _not_built("despeckle_filter")
border_detection_filter = _not_built("border_detection_filter")


def boolean_op(op: int, m1: np.ndarray, m2: np.ndarray, out: np.ndarray) -> None:
    """m[:] = <op>(m1 > 2, m2 > 2) * 255 on mask bodies (uint8, same shape; memmap views welcome)."""
    code = {BOOLEAN_UNION: 0, BOOLEAN_DIFF: 1, BOOLEAN_AND: 2, BOOLEAN_XOR: 3}[op]
    for m in (m1, m2, out):
        if not isinstance(m, np.ndarray) or m.dtype != np.uint8 or m.shape != m1.shape:
            raise TypeError("boolean_op: uint8 masks of one shape expected")
    a, b = dev.to_device(m1 if m1.ndim == 3 else m1[None]), dev.to_device(m2 if m2.ndim == 3 else m2[None])
    o = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.call("b2v_boolean_op", _p(a), _p(b), a.numel(), code, _p(o), _stream())
    dev.to_host(o, out if out.ndim == 3 else out[None])


def convolve_non_zero(volume: np.ndarray, kernel: np.ndarray, cval) -> np.ndarray:
    v = np.ascontiguousarray(volume, dtype=np.float64)
    k = np.ascontiguousarray(kernel, dtype=np.float64)
    if v.ndim != 3 or k.ndim != 3:
        raise TypeError("convolve_non_zero: 3-D float64 volume and kernel expected")
    cval = int(cval)
    if not -32768 <= cval <= 32767:
        raise OverflowError("out of range integral type conversion attempted")    # cval: i16 in the reference
    tv, tk = dev.to_device(v), torch.from_numpy(k).to("cuda")
    out = torch.empty_like(tv)
    with torch.cuda.device(tv.device):
        _lib.call("b2v_convolve_non_zero", _p(tv), *v.shape, _p(tk), *k.shape, float(cval), _p(out), _stream())
    res = np.empty(v.shape, np.float64)
    dev.to_host(out, res)
    return res
