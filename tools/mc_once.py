"""Live timing (CUDA events) of the marching-cubes calls on the bench volume's grown mask."""
import ctypes as C
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scipy.ndimage import generate_binary_structure
from invesalius3_b200 import _lib, device as dev, phantom
from invesalius3_b200.mesh import marching_cubes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = phantom.ct((n, n, n), seed=2)
t = torch.from_numpy(vol).cuda()
seed = phantom.first_seed_in_range(vol, n // 2, 226, 3071)
out = torch.zeros(vol.shape, dtype=torch.uint8, device="cuda")
dev.floodfill_threshold(t, [seed], 226, 3071, 254, generate_binary_structure(3, 1), out)
lib = _lib.load()
ws = dev._workspace(lib.b2v_mc_workspace_bytes(n, n, n), out.device)
nv, nt = C.c_int64(0), C.c_int64(0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tc, te, tw = [], [], []
for it in range(12):
    ev[0].record()
    _lib.call("b2v_mc_count", dev._p(out), _lib.U8, n, n, n, 127.0, dev._p(ws), dev._stream(), C.byref(nv), C.byref(nt))
    ev[1].record()
    verts = torch.empty((nv.value, 3), dtype=torch.float32, device="cuda")
    tris = torch.empty((nt.value, 3), dtype=torch.int32, device="cuda")
    ev[2].record()
    _lib.call("b2v_mc_emit", dev._p(out), _lib.U8, n, n, n, 127.0, dev._p(ws), 1.0, 1.0, 1.0, 0, 0, 0, 1, dev._p(verts),
              dev._p(tris), dev._stream())
    ev[3].record()
    torch.cuda.synchronize()
    tc.append(ev[0].elapsed_time(ev[1])); tw.append(ev[1].elapsed_time(ev[2])); te.append(ev[2].elapsed_time(ev[3]))
print(f"V={nv.value} T={nt.value} count {np.median(tc[2:]):.3f} ms  alloc {np.median(tw[2:]):.3f} ms  emit {np.median(te[2:]):.3f} ms")
ts = []
for it in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); marching_cubes(out, 127, (1, 1, 1), (0, 0, 0), True); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"mesh.marching_cubes {np.median(ts[2:]):.3f} ms (incl. workspace allocation)")
