"""Surface extraction with the reference's piece semantics (numpy in, numpy out).

`contour_piece` reproduces what `create_surface_piece` does up to and including the
contour filter (invesalius/data/surface_process.py:71-186): ROI slicing, optional 1-voxel
border padding (`pad_image`, :52-68), the to_vtk extent/origin arithmetic
(converters.py:34-101), the Y flip about the origin and the iso values (127 on the mask
for Binary / ca_smoothing, tmin and tmax on the image for Default). It returns the mesh as
arrays instead of writing a .vtp; `contour` is the plain array-level entry point.

The pad is applied on the device (a padded copy of the piece), never on the host.

`create_surface_piece` is the reference's 20-argument entry itself (surface_process.py:71-201,
called in spawned worker processes by SurfaceManager, surface.py:1360-1430): memmaps in, the
name of a VTK XML PolyData file (.vtp) out. The file is written without VTK (`write_vtp`: inline
base64 arrays, triangles as Polys), readable by vtkXMLPolyDataReader — what join_process_surface
(surface_process.py:229-268) does next — and by `read_vtp` here.
"""
from __future__ import annotations

import base64
import os
import tempfile
import xml.etree.ElementTree as ET

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
                  index_dtype=np.int32, nz_full: int | None = None):
    """The contour part of create_surface_piece for one Z piece.

    image: int16 [dz][dy][dx] (needed unless from_binary); mask_matrix: the padded uint8
    Mask memmap [dz+1][dy+1][dx+1] (needed when from_binary); roi: slice(z_start, z_stop)
    as built by SurfaceManager.AddNewActor (surface.py:1375-1381, stop may exceed dz)."""
    if from_binary:
        if mask_matrix is None:
            raise ValueError("from_binary needs mask_matrix")
        nz_full = mask_matrix.shape[0] - 1 if nz_full is None else nz_full
        piece = mask_matrix[roi.start + 1:roi.stop + 1, 1:, 1:]
        pad_value, isovalues = 0, [127.0]
    else:
        if image is None:
            raise ValueError("the Default algorithm needs the image")
        nz_full = image.shape[0] if nz_full is None else nz_full
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


# ------------------------------------------------------------------ .vtp (VTK XML PolyData) without VTK
_VTK_TYPES = {np.dtype(np.float32): "Float32", np.dtype(np.int32): "Int32", np.dtype(np.int64): "Int64"}


def _b64(a: np.ndarray) -> str:
    """VTK "binary" DataArray payload: base64(uint32 byte count + raw little-endian data), no compressor."""
    raw = np.ascontiguousarray(a).tobytes()
    return base64.b64encode(np.uint32(len(raw)).tobytes() + raw).decode("ascii")


def write_vtp(filename: str, vertices: np.ndarray, faces: np.ndarray) -> None:
    """Triangle mesh -> VTK XML PolyData (what vtkXMLPolyDataWriter emits for the contour output,
    surface_process.py:188-192: Points + Polys). vertices float32 [V,3], faces int32/int64 [T,3]."""
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    f = np.ascontiguousarray(faces).reshape(-1, 3)
    if f.dtype not in (np.int32, np.int64):
        f = f.astype(np.int64)
    offs = (np.arange(1, f.shape[0] + 1, dtype=f.dtype) * 3)
    it = _VTK_TYPES[f.dtype]
    with open(filename, "w") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="PolyData" version="0.1" byte_order="LittleEndian">\n <PolyData>\n')
        fh.write(f'  <Piece NumberOfPoints="{v.shape[0]}" NumberOfVerts="0" NumberOfLines="0" NumberOfStrips="0" '
                 f'NumberOfPolys="{f.shape[0]}">\n')
        fh.write('   <Points>\n    <DataArray type="Float32" Name="Points" NumberOfComponents="3" format="binary">\n')
        fh.write("     " + _b64(v) + "\n    </DataArray>\n   </Points>\n   <Polys>\n")
        fh.write(f'    <DataArray type="{it}" Name="connectivity" format="binary">\n     ' + _b64(f.reshape(-1)) +
                 "\n    </DataArray>\n")
        fh.write(f'    <DataArray type="{it}" Name="offsets" format="binary">\n     ' + _b64(offs) + "\n    </DataArray>\n")
        fh.write("   </Polys>\n  </Piece>\n </PolyData>\n</VTKFile>\n")


def read_vtp(filename: str):
    """Inverse of write_vtp (inline base64, uncompressed, UInt32 headers): (vertices, faces)."""
    root = ET.parse(filename).getroot()
    piece = root.find("PolyData").find("Piece")

    def arr(node):
        dt = {v: k for k, v in _VTK_TYPES.items()}[node.get("type")]
        raw = base64.b64decode(node.text.strip())
        n = int(np.frombuffer(raw[:4], np.uint32)[0])
        return np.frombuffer(raw[4:4 + n], dt).copy()

    pts = arr(piece.find("Points").find("DataArray")).reshape(-1, 3)
    polys = {d.get("Name"): arr(d) for d in piece.find("Polys").findall("DataArray")}
    conn = polys["connectivity"].reshape(-1, 3)
    assert int(piece.get("NumberOfPoints")) == pts.shape[0] and int(piece.get("NumberOfPolys")) == conn.shape[0]
    return pts, conn


def create_surface_piece(filename, shape, dtype, mask_filename, mask_shape, mask_dtype, roi, spacing, mode, min_value,
                         max_value, decimate_reduction, smooth_relaxation_factor, smooth_iterations, language,
                         flip_image, from_binary, algorithm, imagedata_resolution, fill_border_holes):
    """invesalius/data/surface_process.py:71-201 with its own signature and result: the two memmaps in
    (image `filename`, padded mask `mask_filename`), the name of the written .vtp piece out. Runs in
    a spawned worker process like the reference's (surface.py:1368-1369): everything it needs is
    imported here, CUDA initialises on first use. The arguments the reference's body ignores as well
    (mode, decimate_reduction, smooth_*, language, flip_image, imagedata_resolution) are accepted and
    ignored. algorithm "InVesalius 3.b2" (vtkImageGaussianSmooth before the contour) is not built."""
    if not from_binary and algorithm == "InVesalius 3.b2":
        raise NotImplementedError("create_surface_piece: the 'InVesalius 3.b2' pre-smoothing is not built on the device")
    mask = np.memmap(mask_filename, mode="r", dtype=mask_dtype, shape=tuple(mask_shape))
    image = None if from_binary else np.memmap(filename, mode="r", dtype=dtype, shape=tuple(shape))
    # contour_piece derives pad_top from the array it is given; the reference uses `shape` (the image's)
    verts, faces = contour_piece(image, mask if from_binary else None, roi, spacing, min_value, max_value,
                                 from_binary=bool(from_binary), fill_border_holes=bool(fill_border_holes), flip_y=True,
                                 index_dtype=np.int64, nz_full=int(shape[0]))
    fd, out = tempfile.mkstemp(suffix="_%d_%d.vtp" % (roi.start, roi.stop))
    os.close(fd)
    write_vtp(out, verts, faces)
    return out
