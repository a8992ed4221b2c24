"""Three steps of the headline pipeline (threshold -> flood fill -> marching cubes on the 512^3 phantom),
for ncu captures: `ncu -k regex:k_ -s <2 steps of launches> -c <1 step>`."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scipy.ndimage import generate_binary_structure
from invesalius3_b200 import _lib, device as dev, phantom
from invesalius3_b200.mesh import marching_cubes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
vol = phantom.ct((n, n, n), seed=2)
t = torch.from_numpy(vol).cuda()
seed = phantom.first_seed_in_range(vol, n // 2, 226, 3071)
st = generate_binary_structure(3, 1)
out = torch.empty(vol.shape, dtype=torch.uint8, device="cuda")
mask = torch.empty(vol.shape, dtype=torch.uint8, device="cuda")
lib = _lib.load()
for it in range(steps):
    lib.b2v_launch_count_reset()
    dev.threshold(t, 226, 3071, out=mask)
    out.zero_()
    dev.floodfill_threshold(t, [seed], 226, 3071, 254, st, out)
    v, f = marching_cubes(out, 127, (1, 1, 1), (0, 0, 0), True)
    torch.cuda.synchronize()
print("launches per step", lib.b2v_launch_count(), "V", v.shape[0], "T", f.shape[0])
