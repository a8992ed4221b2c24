"""Where does the numpy-in/numpy-out time go? Times each transfer / kernel piece of the
three reference-shaped calls on a 512^3 volume with pinned host buffers."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from invesalius3_b200 import device as dev, phantom  # noqa: E402


def t(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    n = 512
    vol = phantom.ct((n, n, n), seed=2)
    h_vol = torch.from_numpy(vol).pin_memory().numpy()
    h_mask = torch.zeros((n + 1, n + 1, n + 1), dtype=torch.uint8).pin_memory().numpy()
    h_out = torch.zeros((n, n, n), dtype=torch.uint8).pin_memory().numpy()
    pageable = np.zeros((n, n, n), np.uint8)
    d_vol = dev.to_device(h_vol)
    d_mask = dev.threshold(d_vol, 226, 3071)
    gb = lambda nbytes, ms: nbytes / ms / 1e6
    r = {}
    ms = t(lambda: dev.to_device(h_vol)); r["H2D image 256MiB pinned dense"] = (ms, gb(vol.nbytes, ms))
    ms = t(lambda: dev.to_device(vol)); r["H2D image 256MiB pageable"] = (ms, gb(vol.nbytes, ms))
    ms = t(lambda: dev.to_host(d_mask, h_out)); r["D2H mask 128MiB pinned dense"] = (ms, gb(h_out.nbytes, ms))
    ms = t(lambda: dev.to_host(d_mask, pageable)); r["D2H mask 128MiB pageable dense"] = (ms, gb(h_out.nbytes, ms))
    ms = t(lambda: dev.to_host(d_mask, h_mask[1:, 1:, 1:])); r["D2H mask 128MiB pinned strided [1:,1:,1:]"] = (ms, gb(h_out.nbytes, ms))
    ms = t(lambda: dev.to_device(h_mask[1:, 1:, 1:])); r["H2D mask 128MiB pinned strided"] = (ms, gb(h_out.nbytes, ms))
    ms = t(lambda: h_out.fill(0)); r["host memset 128MiB"] = (ms, gb(h_out.nbytes, ms))
    ms = t(lambda: d_mask.cpu().numpy()); r["tensor.cpu() 128MiB (pageable alloc)"] = (ms, gb(h_out.nbytes, ms))
    # the three reference-shaped calls of bench.py's e2e leg
    from scipy.ndimage import generate_binary_structure
    from invesalius3_b200 import invesalius_rs, slice_ops, surface_process
    seed = phantom.first_seed_in_range(vol, n // 2, 226, 3071)
    st = generate_binary_structure(3, 1)
    th_out = torch.zeros((n, n, n), dtype=torch.uint8).pin_memory()
    ms = t(lambda: slice_ops.set_mask_threshold(h_vol, h_mask, (226, 3071))); r["call: set_mask_threshold"] = (ms, 0)
    ms = t(lambda: th_out.zero_()); r["call: zero the out mask (torch, pinned)"] = (ms, gb(h_out.nbytes, ms))
    np_o = th_out.numpy()
    ms = t(lambda: invesalius_rs.floodfill_threshold(h_vol, [seed], 226, 3071, 254, st, np_o)); r["call: floodfill_threshold"] = (ms, 0)
    ms = t(lambda: surface_process.contour(np_o, [127], (1, 1, 1), 0, True)); r["call: contour"] = (ms, 0)
    for k, (ms, g) in r.items():
        print(f"{k:48s} {ms:8.2f} ms  {g:7.1f} GB/s")


if __name__ == "__main__":
    main()
