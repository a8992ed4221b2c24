"""Step through the hot path on a small phantom with a synchronise after every op (debug aid)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scipy.ndimage import generate_binary_structure
from invesalius3_b200 import device, phantom
from invesalius3_b200.mesh import marching_cubes
import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
vol = phantom.ct((n, n, n), seed=2)
t = torch.from_numpy(vol).cuda()
m = device.threshold(t, 226, 3071); torch.cuda.synchronize(); print("threshold ok", flush=True)
st = generate_binary_structure(3, 1)
seed = phantom.first_seed_in_range(vol, n // 2, 226, 3071)
out = torch.zeros(vol.shape, dtype=torch.uint8, device="cuda")
r = device.floodfill_threshold(t, [seed], 226, 3071, 254, st, out); torch.cuda.synchronize(); print("flood ok", r, flush=True)
ref = np.zeros(vol.shape, np.uint8)
oracle.floodfill_threshold(vol, [seed], 226, 3071, 254, st, ref)
print("flood equal", np.array_equal(out.cpu().numpy(), ref), flush=True)
v, f = marching_cubes(out, 127, (1, 1, 1), (0, 0, 0), True); torch.cuda.synchronize(); print("mc ok", v.shape, f.shape, flush=True)
vo, fo = oracle.marching_cubes(ref, 127, (1, 1, 1), (0, 0, 0), True)
print("mc counts", vo.shape, fo.shape, flush=True)
print("tris equal", np.array_equal(f.cpu().numpy(), fo), "verts", float(np.abs(v.cpu().numpy() - vo).max()) if len(vo) == len(v) else None, flush=True)
