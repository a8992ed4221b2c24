"""One small, odd-shaped invocation of every kernel family (run under compute-sanitizer:
`compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scipy.ndimage import generate_binary_structure  # noqa: E402

from invesalius3_b200 import _lib, device as dev, invesalius_rs as rs, phantom, slice_ops, surface_process  # noqa: E402
from invesalius3_b200 import watershed_process as wp  # noqa: E402

lib = _lib.load()
for shape in [(7, 9, 11), (13, 21, 70), (24, 40, 56)]:
    vol = phantom.ct(shape, seed=3)
    t = torch.from_numpy(vol).cuda()
    m = dev.threshold(t, 226, 3071)
    dev.threshold(t, 226, 3071, out=m, preserve_markers=True)
    mm = np.zeros(tuple(s + 1 for s in shape), np.uint8)
    slice_ops.do_threshold_to_all_slices(vol, mm, (226, 3071))
    for axis in (0, 1, 2):
        for kind in ("max", "min", "mean"):
            dev.mip(t, axis, kind)
        o = np.zeros([(shape[1], shape[2]), (shape[0], shape[2]), (shape[0], shape[1])][axis], np.int16)
        rs.mida(vol, axis, 300, 600, o)
        rs.lmip(vol, axis, 700, 3033, o)
        for tmip in (0, 1, 2):
            rs.fast_countour_mip(vol, 1.5, axis, 300, 300, tmip, o)
    for eng in (1, 0):
        lib.b2v_floodfill_set_engine(eng)
        for conn in (1, 2, 3):
            out = np.zeros(shape, np.uint8)
            seed = phantom.first_seed_in_range(vol, shape[0] // 2, -2000, 4000)
            rs.floodfill_threshold(vol, [seed, (0, 0, 0)], -200, 3071, 254, generate_binary_structure(3, conn), out)
        rs.floodfill_threshold_inplace(mm[1:, 1:, 1:], [(1, 1, 1)], 0, 255, 7, generate_binary_structure(3, 1))
        rs.floodfill(vol, 1, 1, 1, int(vol[1, 1, 1]), 9, np.zeros(shape, np.uint8))
    lib.b2v_floodfill_set_engine(1)
    lab = np.arange(vol.size, dtype=np.uint32).reshape(shape) % 5
    rs.fill_holes_automatically(np.zeros(shape, np.uint8), lab, 4, 10 ** 6)
    surface_process.contour((vol > 100).astype(np.uint8) * 255, [127], (0.5, 0.5, 1.0))
    surface_process.contour(vol, [226, 3071], (0.5, 0.5, 1.0))
    mk = np.zeros(shape, np.uint8); mk[1, 1, 1] = 1; mk[-2, -2, -2] = 2
    for alg in ("Watershed", "Watershed IFT"):
        wp.watershed_device(t, torch.from_numpy(mk).cuda(), generate_binary_structure(3, 1), alg, 3, True, -18, 406)
torch.cuda.synchronize()
print("sanitize smoke ok,", lib.b2v_launch_count(), "launches")
