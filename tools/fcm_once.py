"""Contour-MIP timings at 1024^3 (CUDA events): the contour volume alone and the whole call."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from invesalius3_b200 import projection
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator(device="cuda").manual_seed(0)
vol = torch.randint(-1024, 3072, (n, n, n), dtype=torch.int16, device="cuda", generator=g)
N = vol.numel()
def t(fn, it=5):
    fn(); fn()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for nn in (1.0, 2.0):
    for axis in (0, 2):
        ms = t(lambda: projection.fcm_volume(vol, nn, axis))
        print(f"fcm_volume n={nn} axis={axis}: {ms:.3f} ms  {4 * N / ms / 1e6:.0f} GB/s moved (2 B read + 2 B written)")
for tmip in (0, 1, 2):
    ms = t(lambda: projection.fast_countour_mip(vol, 2.0, 0, 300, 300, tmip))
    print(f"fast_countour_mip tmip={tmip} axis=0: {ms:.3f} ms  {2 * N / ms / 1e6:.0f} GB/s at 2 B/voxel")
