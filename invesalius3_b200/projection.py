"""MIDA / LMIP / contour-MIP on device tensors (C ABI: b2v_mida, b2v_lmip,
b2v_fast_countour_mip). Each call synchronises once (error status of the casts)."""
from __future__ import annotations

import torch

from . import _lib
from .device import _dense, _p, _stream, _workspace, dtype_code

_OUT_SHAPE = lambda s, axis: [(s[1], s[2]), (s[0], s[2]), (s[0], s[1])][axis]  # noqa: E731


def _prep(image: torch.Tensor, axis: int, out: torch.Tensor | None, out_dtype):
    _dense(image, "image")
    if image.dim() != 3:
        raise TypeError("image must be 3-dimensional")
    if axis not in (0, 1, 2):
        raise ValueError("axis must be 0, 1 or 2")
    shape = _OUT_SHAPE(image.shape, axis)
    if out is None:
        out = torch.empty(shape, dtype=out_dtype, device=image.device)
    _dense(out, "out")
    if tuple(out.shape) != shape:
        raise ValueError(f"out must have shape {shape}")
    ws = _workspace(_lib.load().b2v_proj_workspace_bytes(image.numel()), image.device)
    return out, ws


def mida(image: torch.Tensor, axis: int, wl, ww, out: torch.Tensor | None = None,
         minmax: torch.Tensor | None = None) -> torch.Tensor:
    """mips.rs:102-168. wl/ww are interpreted as the image dtype, like the PyO3 layer.
    `minmax` (float32 [2] on the device) replaces the pass over the volume that finds them —
    a Z shard passes the all-reduced global pair."""
    odt = {torch.int16: torch.int16, torch.uint8: torch.uint8, torch.float64: torch.uint8}.get(image.dtype)
    if odt is None:
        raise TypeError("Invalid image or output type")
    out, ws = _prep(image, axis, out, odt)
    if out.dtype not in (torch.int16, torch.uint8):
        raise TypeError("Invalid image or output type")
    dz, dy, dx = image.shape
    with torch.cuda.device(image.device):
        if minmax is None:
            _lib.call("b2v_mida", _p(image), dtype_code(image), dz, dy, dx, axis, float(wl), float(ww), _p(out),
                      dtype_code(out), _p(ws), _stream())
        else:
            if minmax.dtype != torch.float32 or minmax.numel() != 2 or not minmax.is_cuda:
                raise TypeError("minmax must be a float32 CUDA tensor with 2 elements")
            _lib.call("b2v_mida_minmax", _p(image), dtype_code(image), dz, dy, dx, axis, float(wl), float(ww),
                      _p(minmax.contiguous()), _p(out), dtype_code(out), _p(ws), _stream())
    return out


def lmip(image: torch.Tensor, axis: int, tmin, tmax, out: torch.Tensor | None = None) -> torch.Tensor:
    """mips.rs:7-86."""
    out, ws = _prep(image, axis, out, image.dtype)
    if out.dtype != image.dtype:
        raise TypeError("Invalid image or output type")
    dz, dy, dx = image.shape
    with torch.cuda.device(image.device):
        _lib.call("b2v_lmip", _p(image), dtype_code(image), dz, dy, dx, axis, float(tmin), float(tmax), _p(out), _p(ws),
                  _stream())
    return out


def fcm_volume(image: torch.Tensor, n: float, axis: int) -> torch.Tensor:
    """The contour volume tmp[z, y, x] = T(calc_fcm_intensity(image, x, y, z, n, dir(axis)))
    (mips.rs:197-242), same dtype as the image (int16, uint8 or float64)."""
    _dense(image, "image")
    if image.dim() != 3:
        raise TypeError("image must be 3-dimensional")
    if axis not in (0, 1, 2):
        raise ValueError("axis must be 0, 1 or 2")
    tmp = torch.empty_like(image)
    ws = _workspace(_lib.load().b2v_proj_workspace_bytes(image.numel()), image.device)
    dz, dy, dx = image.shape
    with torch.cuda.device(image.device):
        _lib.call("b2v_fcm_volume", _p(image), dtype_code(image), dz, dy, dx, float(n), axis, _p(tmp), _p(ws), _stream())
    return tmp


def fast_countour_mip(image: torch.Tensor, n: float, axis: int, wl, ww, tmip: int,
                      out: torch.Tensor | None = None) -> torch.Tensor:
    """mips.rs:215-279 (tmip 0 max, 1 LMIP(700, 3033), 2 MIDA): contour volume, then its projection."""
    out, _ = _prep(image, axis, out, image.dtype)
    if out.dtype != image.dtype:
        raise TypeError("Invalid image or output type")
    dz, dy, dx = image.shape
    code = dtype_code(image)
    ws = _workspace(_lib.load().b2v_fcm_workspace_bytes(code, dz, dy, dx, axis, int(tmip)), image.device)
    with torch.cuda.device(image.device):
        _lib.call("b2v_fast_countour_mip", _p(image), code, dz, dy, dx, float(n), axis, float(wl),
                  float(ww), int(tmip), _p(out), _p(ws), _stream())
    return out


# ---- rays along z over one Z shard (dist.mida / dist.lmip, axis 0) ---------------------------
def ray_state(image: torch.Tensor) -> torch.Tensor:
    """Per-ray state handed from shard to shard: uint32 words [3][dy][dx] (b2v.h)."""
    return torch.zeros((3, image.shape[1], image.shape[2]), dtype=torch.int32, device=image.device)


def _z_partial_prep(image, state, last, out_dtype):
    _dense(image, "image")
    if image.dim() != 3:
        raise TypeError("image must be 3-dimensional")
    if state.dtype != torch.int32 or tuple(state.shape) != (3, image.shape[1], image.shape[2]) or not state.is_cuda:
        raise TypeError("state must be an int32 CUDA tensor of shape (3, dy, dx)")
    _dense(state, "state")
    out = torch.empty((image.shape[1], image.shape[2]), dtype=out_dtype, device=image.device) if last else None
    ws = _workspace(_lib.load().b2v_proj_workspace_bytes(image.numel()), image.device)
    return out, ws


def mida_z_partial(image: torch.Tensor, wl, ww, minmax: torch.Tensor, state: torch.Tensor, first: bool,
                   last: bool) -> torch.Tensor | None:
    """One shard's stretch of the MIDA rays along z (b2v_mida_z_partial). Returns the image on
    the last shard, None elsewhere; `state` is updated in place."""
    odt = {torch.int16: torch.int16, torch.uint8: torch.uint8, torch.float64: torch.uint8}.get(image.dtype)
    if odt is None:
        raise TypeError("Invalid image or output type")
    if minmax.dtype != torch.float32 or minmax.numel() != 2 or not minmax.is_cuda:
        raise TypeError("minmax must be a float32 CUDA tensor with 2 elements")
    out, ws = _z_partial_prep(image, state, last, odt)
    dz, dy, dx = image.shape
    with torch.cuda.device(image.device):
        _lib.call("b2v_mida_z_partial", _p(image), dtype_code(image), dz, dy, dx, float(wl), float(ww),
                  _p(minmax.contiguous()), _p(state), int(bool(first)), int(bool(last)),
                  _p(out) if out is not None else None, _lib.I16 if odt == torch.int16 else _lib.U8, _p(ws),
                  _stream())
    return out


def lmip_z_partial(image: torch.Tensor, tmin, tmax, state: torch.Tensor, first: bool, last: bool):
    """One shard's stretch of the LMIP rays along z (b2v_lmip_z_partial)."""
    out, ws = _z_partial_prep(image, state, last, image.dtype)
    dz, dy, dx = image.shape
    with torch.cuda.device(image.device):
        _lib.call("b2v_lmip_z_partial", _p(image), dtype_code(image), dz, dy, dx, float(tmin), float(tmax),
                  _p(state), int(bool(first)), int(bool(last)), _p(out) if out is not None else None, _p(ws),
                  _stream())
    return out

