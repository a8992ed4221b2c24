"""MIDA / LMIP with rays along x at 1024^3: TMA-staged rows (B2V_TMA=1) against the lane-load kernels (default)."""
import os, sys
from pathlib import Path
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
tag = "TMA rows" if os.environ.get("B2V_TMA") else "lane loads"
for axis in (2, 0):
    ms = t(lambda: projection.mida(vol, axis, 32000, 2))
    print(f"[{tag}] MIDA full rays axis {axis}: {ms:.3f} ms  {4 * N / ms / 1e6:.0f} GB/s (4 B/voxel)")
    ms = t(lambda: projection.mida(vol, axis, 300, 300))
    print(f"[{tag}] MIDA early exit axis {axis}: {ms:.3f} ms")
    ms = t(lambda: projection.lmip(vol, axis, 700, 3033))
    print(f"[{tag}] LMIP axis {axis}: {ms:.3f} ms")
