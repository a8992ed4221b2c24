"""Marching cubes on device tensors (C ABI: b2v_mc_count / b2v_mc_emit)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .device import _dense, _p, _stream, _workspace, dtype_code


def marching_cubes(vol: torch.Tensor, iso: float, spacing=(1.0, 1.0, 1.0), origin_index=(0, 0, 0),
                   flip_y: bool = True, _events: list | None = None):
    """Iso-surface of a dense uint8/int16 [nz][ny][nx] device volume.

    spacing = (sx, sy, sz); origin_index = (ox, oy, oz) is added to the (x, y, z) voxel
    indices before scaling (padding / piece offset, converters.py:55-63); flip_y negates y
    (vtkImageFlip about the origin, surface_process.py:156-161).
    Returns (vertices float32 [V,3], triangles int32 [T,3]) on the device; vertices are
    shared between triangles. Synchronises once (the counts come back to the host)."""
    _dense(vol, "vol")
    if vol.dim() != 3:
        raise ValueError("marching_cubes: 3-D volume expected")
    code = dtype_code(vol)
    if code not in (_lib.I16, _lib.U8):
        raise TypeError("marching_cubes: volume must be uint8 or int16")
    nz, ny, nx = vol.shape
    lib = _lib.load()
    ws = _workspace(lib.b2v_mc_workspace_bytes(nz, ny, nx), vol.device)
    nv, nt = C.c_int64(0), C.c_int64(0)
    with torch.cuda.device(vol.device):
        if _events is not None: _events[0].record()
        _lib.call("b2v_mc_count", _p(vol), code, nz, ny, nx, float(iso), _p(ws), _stream(), C.byref(nv), C.byref(nt))
        if _events is not None: _events[1].record()
        verts = torch.empty((nv.value, 3), dtype=torch.float32, device=vol.device)
        tris = torch.empty((nt.value, 3), dtype=torch.int32, device=vol.device)
        if _events is not None: _events[2].record()
        if nv.value or nt.value:
            _lib.call("b2v_mc_emit", _p(vol), code, nz, ny, nx, float(iso), _p(ws), float(spacing[0]),
                      float(spacing[1]), float(spacing[2]), int(origin_index[0]), int(origin_index[1]),
                      int(origin_index[2]), int(bool(flip_y)), _p(verts), _p(tris), _stream())
    return verts, tris
