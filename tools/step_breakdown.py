"""Where the bench step's time goes: the three stages with the MC stage split into count (GPU +
host sync), output allocation and emit; host wall clock beside the CUDA events."""
import ctypes as C
import sys
import time
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scipy.ndimage import generate_binary_structure
from invesalius3_b200 import _lib, device as dev, phantom
from invesalius3_b200.mesh import marching_cubes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mode = sys.argv[2] if len(sys.argv) > 2 else "split"
vol = phantom.ct((n, n, n), seed=2)
t = torch.from_numpy(vol).cuda()
seed = phantom.first_seed_in_range(vol, n // 2, 226, 3071)
st = generate_binary_structure(3, 1)
out = torch.empty(vol.shape, dtype=torch.uint8, device="cuda")
mask = torch.empty(vol.shape, dtype=torch.uint8, device="cuda")
lib = _lib.load()
nv, nt = C.c_int64(0), C.c_int64(0)
rows = []
keep = None
for it in range(14):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    h = [time.perf_counter()]
    ev[0].record()
    dev.threshold(t, 226, 3071, out=mask)
    ev[1].record(); h.append(time.perf_counter())
    out.zero_()
    dev.floodfill_threshold(t, [seed], 226, 3071, 254, st, out)
    ev[2].record(); h.append(time.perf_counter())
    if mode == "split":
        ws = dev._workspace(lib.b2v_mc_workspace_bytes(n, n, n), out.device)
        ev[3].record(); h.append(time.perf_counter())
        _lib.call("b2v_mc_count", dev._p(out), _lib.U8, n, n, n, 127.0, dev._p(ws), dev._stream(), C.byref(nv), C.byref(nt))
        ev[4].record(); h.append(time.perf_counter())
        verts = torch.empty((nv.value, 3), dtype=torch.float32, device="cuda")
        tris = torch.empty((nt.value, 3), dtype=torch.int32, device="cuda")
        ev[5].record(); h.append(time.perf_counter())
        _lib.call("b2v_mc_emit", dev._p(out), _lib.U8, n, n, n, 127.0, dev._p(ws), 1.0, 1.0, 1.0, 0, 0, 0, 1, dev._p(verts),
                  dev._p(tris), dev._stream())
        ev[6].record(); h.append(time.perf_counter())
    else:
        for k in (3, 4, 5):
            ev[k].record(); h.append(time.perf_counter())
        verts, tris = marching_cubes(out, 127, (1, 1, 1), (0, 0, 0), True)
        ev[6].record(); h.append(time.perf_counter())
    torch.cuda.synchronize()
    keep = (verts, tris)
    rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(6)] + [(h[i + 1] - h[i]) * 1e3 for i in range(6)])
r = np.median(np.array(rows[4:]), axis=0)
names = ["threshold", "zero+flood", "mc ws alloc", "mc count", "mc out alloc", "mc emit"]
print(f"n={n} mode={mode} V={nv.value} T={nt.value}")
for i, k in enumerate(names):
    print(f"  {k:14s} gpu {r[i]:.3f} ms   host {r[6 + i]:.3f} ms")
