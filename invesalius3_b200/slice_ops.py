"""Threshold entry points with the reference's numpy-in / numpy-out signatures.

Drop-in replacements for the per-voxel statements of `invesalius.data.slice_.Slice`:
  do_threshold_to_a_slice      slice_.py:1722-1737
  set_mask_threshold           slice_.py:1238-1246 (whole-volume branch of SetMaskThreshold)
  set_mask_threshold_slice     slice_.py:1252-1256 (per-slice preview branch)
  do_threshold_to_all_slices   slice_.py:1739-1769
The `Slice` methods keep their bookkeeping (Project lookup, pubsub); a maintainer swaps
the NumPy bodies for these calls.

`mask_matrix` is the padded uint8 memmap of `Mask` (shape (dz+1, dy+1, dx+1), flags in
[n,0,0]; invesalius/data/mask.py:422-431, :462-466).
"""
from __future__ import annotations

import numpy as np
import torch

from . import device as dev


def _check_image(a: np.ndarray, ndim: int) -> None:
    if not isinstance(a, np.ndarray) or a.ndim != ndim:
        raise TypeError(f"expected a {ndim}-D numpy array")
    if a.dtype != np.int16:
        raise TypeError(f"threshold: image must be int16, got {a.dtype}")


def do_threshold_to_a_slice(slice_matrix: np.ndarray, mask: np.ndarray, threshold) -> np.ndarray:
    """((v>=tmin)&(v<=tmax))*255 with old-mask markers 1/2/253/254 kept; returns uint8."""
    _check_image(slice_matrix, 2)
    if mask.shape != slice_matrix.shape:
        raise ValueError("mask and slice shapes differ")
    tmin, tmax = threshold
    img = dev.to_device(slice_matrix[None])
    old = dev.to_device(np.asarray(mask, dtype=np.uint8)[None])
    dev.threshold(img, tmin, tmax, out=old, preserve_markers=True)
    res = np.empty(slice_matrix.shape, dtype=np.uint8)
    dev.to_host(old, res[None])
    return res


def set_mask_threshold_slice(slice_image: np.ndarray, threshold_range) -> np.ndarray:
    """(255 * ((img >= tmin) & (img <= tmax))).astype('uint8') for one buffered slice."""
    _check_image(slice_image, 2)
    tmin, tmax = threshold_range
    out = dev.threshold(dev.to_device(slice_image[None]), tmin, tmax)
    res = np.empty(slice_image.shape, dtype=np.uint8)
    dev.to_host(out, res[None])
    return res


def set_mask_threshold(matrix: np.ndarray, mask_matrix: np.ndarray, threshold_range) -> None:
    """Whole-volume SetMaskThreshold: every slice rewritten (no marker preservation),
    axial flags mask_matrix[n+1,0,0] = 1."""
    _check_image(matrix, 3)
    dz, dy, dx = matrix.shape
    if mask_matrix.shape != (dz + 1, dy + 1, dx + 1) or mask_matrix.dtype != np.uint8:
        raise ValueError("mask_matrix must be uint8 of shape (dz+1, dy+1, dx+1)")
    tmin, tmax = threshold_range
    body = mask_matrix[1:, 1:, 1:]
    if dz >= 32 and dev._box_pitches(matrix) is not None and dev._box_pitches(body) is not None:
        _threshold_pipelined(matrix, body, tmin, tmax)
    else:
        out = dev.threshold(dev.to_device(matrix), tmin, tmax)
        dev.to_host(out, body)
    mask_matrix[1:, 0, 0] = 1


def _threshold_pipelined(matrix: np.ndarray, body: np.ndarray, tmin, tmax, nchunks: int = 4) -> None:
    """Z slabs on two streams: slab k's device->host copy runs while slab k+1 travels host->
    device (PCIe is full duplex), with the threshold kernel in between. Pays off with pinned
    host memory; with pageable memory the copies are staged by the driver and merely
    serialise, the result being the same."""
    dev.require_cuda()
    dz = matrix.shape[0]
    img = torch.empty(matrix.shape, dtype=torch.int16, device="cuda")
    out = torch.empty(matrix.shape, dtype=torch.uint8, device="cuda")
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for s in streams:
        s.wait_stream(cur)
    bounds = [dz * k // nchunks for k in range(nchunks + 1)]
    for k in range(nchunks):
        z0, z1 = bounds[k], bounds[k + 1]
        with torch.cuda.stream(streams[k % 2]):
            dev.h2d_async(img[z0:z1], matrix[z0:z1])
            dev.threshold(img[z0:z1], tmin, tmax, out=out[z0:z1])
            dev.d2h_async(out[z0:z1], body[z0:z1])
    for s in streams:
        s.synchronize()
        cur.wait_stream(s)


def do_threshold_to_all_slices(matrix: np.ndarray, mask_matrix: np.ndarray, threshold_range) -> None:
    """Threshold every axial slice whose flag mask_matrix[n,0,0] is 0, keeping markers,
    then set the flag (slice_.py:1762-1767). Flushes memmaps like the reference."""
    _check_image(matrix, 3)
    dz, dy, dx = matrix.shape
    if mask_matrix.shape != (dz + 1, dy + 1, dx + 1) or mask_matrix.dtype != np.uint8:
        raise ValueError("mask_matrix must be uint8 of shape (dz+1, dy+1, dx+1)")
    tmin, tmax = threshold_range
    flags = np.asarray(mask_matrix[1:, 0, 0])
    dirty = np.flatnonzero(flags == 0)
    if dirty.size:
        # contiguous runs of dirty slices go to the device as slabs
        runs = np.split(dirty, np.flatnonzero(np.diff(dirty) != 1) + 1)
        for run in runs:
            z0, z1 = int(run[0]), int(run[-1]) + 1
            img = dev.to_device(matrix[z0:z1])
            view = mask_matrix[z0 + 1:z1 + 1, 1:, 1:]
            old = dev.to_device(view)
            dev.threshold(img, tmin, tmax, out=old, preserve_markers=True)
            dev.to_host(old, view)
        mask_matrix[dirty + 1, 0, 0] = 1
    if hasattr(mask_matrix, "flush"):
        mask_matrix.flush()
