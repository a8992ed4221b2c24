"""Surface extraction with the reference's piece semantics (numpy in, numpy out).

`contour_piece` reproduces what `create_surface_piece` does up to and including the
contour filter (invesalius/data/surface_process.py:71-186): ROI slicing, optional 1-voxel
border padding (`pad_image`, :52-68), the to_vtk extent/origin arithmetic
(converters.py:34-101), the Y flip about the origin and the iso values (127 on the mask
for Binary / ca_smoothing, tmin and tmax on the image for Default). It returns the mesh as
arrays instead of writing a .vtp; `contour` is the plain array-level entry point.

The pad is applied on the device (a padded copy of the piece), never on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import device as dev
from .mesh import marching_cubes


def contour(volume: np.ndarray, isovalues, spacing=(1.0, 1.0, 1.0), z0: int = 0, flip_y: bool = True,
            padding=(0, 0, 0), index_dtype=np.int32):
    """Iso-surfaces of `volume` (uint8 or int16, [z][y][x]) at each value of `isovalues`.

    spacing = (sx, sy, sz); z0 = index of the first slice in the full volume; padding =
    (px, py, pz) voxels already added in front of the data (subtracted from the indices as
    in converters.to_vtk). Returns (vertices float32 [V,3], faces [T,3] of `index_dtype`);
    the surfaces of successive isovalues are concatenated in order. int32 faces (what
    `invesalius_rs.Mesh` takes as FaceArray::I32, types.rs:63-70) halve the device->host
    traffic; pass index_dtype=np.int64 for vtkIdType-sized indices."""
    if not isinstance(volume, np.ndarray) or volume.ndim != 3:
        raise TypeError("contour: 3-D numpy volume expected")
    if volume.dtype not in (np.uint8, np.int16):
        raise TypeError("contour: volume must be uint8 or int16")
    isovalues = [float(v) for v in np.atleast_1d(isovalues)]
    t = dev.to_device(volume)
    return _contour_device(t, isovalues, spacing, z0, flip_y, padding, index_dtype)


def _contour_device(t: torch.Tensor, isovalues, spacing, z0, flip_y, padding, index_dtype=np.int32):
    px, py, pz = padding
    tdt = torch.int64 if np.dtype(index_dtype) == np.int64 else torch.int32
    vs, fs, base = [], [], 0
    for iso in isovalues:
        v, f = marching_cubes(t, iso, spacing, (-px, -py, z0 - pz), flip_y)
        vs.append(v)
        f = f.to(tdt)
        fs.append(f + base if base else f)
        base += v.shape[0]
    verts = torch.cat(vs) if len(vs) > 1 else vs[0]
    faces = torch.cat(fs) if len(fs) > 1 else fs[0]
    return dev.to_numpy(verts), dev.to_numpy(faces)


def _pad_device(t: torch.Tensor, pad_value: int, pad_bottom: bool, pad_top: bool) -> torch.Tensor:
    """pad_image (surface_process.py:52-68) on the device."""
    dz, dy, dx = t.shape
    z_iadd = 1 if pad_bottom else 0
    out = torch.full((dz + z_iadd + (1 if pad_top else 0), dy + 2, dx + 2), pad_value, dtype=t.dtype,
                     device=t.device)
    out[z_iadd:z_iadd + dz, 1:-1, 1:-1] = t
    return out


def contour_piece(image: np.ndarray | None, mask_matrix: np.ndarray | None, roi: slice, spacing, min_value=None,
                  max_value=None, from_binary: bool = True, fill_border_holes: bool = True, flip_y: bool = True,
                  index_dtype=np.int32):
    """The contour part of create_surface_piece for one Z piece.

    image: int16 [dz][dy][dx] (needed unless from_binary); mask_matrix: the padded uint8
    Mask memmap [dz+1][dy+1][dx+1] (needed when from_binary); roi: slice(z_start, z_stop)
    as built by SurfaceManager.AddNewActor (surface.py:1375-1381, stop may exceed dz)."""
    if from_binary:
        if mask_matrix is None:
            raise ValueError("from_binary needs mask_matrix")
        nz_full = mask_matrix.shape[0] - 1
        piece = mask_matrix[roi.start + 1:roi.stop + 1, 1:, 1:]
        pad_value, isovalues = 0, [127.0]
    else:
        if image is None:
            raise ValueError("the Default algorithm needs the image")
        nz_full = image.shape[0]
        piece = image[roi]
        pad_value, isovalues = int(np.iinfo(image.dtype).min), [float(min_value), float(max_value)]
    if piece.shape[0] == 0:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), index_dtype)
    pad_bottom = roi.start == 0
    pad_top = roi.stop >= nz_full
    t = dev.to_device(piece)
    if fill_border_holes:
        t = _pad_device(t, pad_value, pad_bottom, pad_top)
        padding = (1, 1, int(pad_bottom))
    else:
        padding = (0, 0, 0)
    return _contour_device(t, isovalues, spacing, roi.start, flip_y, padding, index_dtype)
