"""One 512^3 watershed per algorithm (for ncu launch lists)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from scipy.ndimage import generate_binary_structure
from invesalius3_b200 import phantom, watershed_process as wp
from ws_bench import markers_for
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = phantom.ct((n, n, n), seed=4)
mk = markers_for(vol, 4)
st = generate_binary_structure(3, 1)
t_vol, t_mk = torch.from_numpy(vol).cuda(), torch.from_numpy(mk).cuda()
for alg in ("Watershed", "Watershed IFT"):
    wp.watershed_device(t_vol, t_mk, st, alg, 3, True, -18, 406)
    torch.cuda.synchronize()
    import ctypes as C
    from invesalius3_b200 import _lib
    st8 = (C.c_int * 8)()
    _lib.load().b2v_ws_stats(st8, 1)
    print(alg, "rounds", wp.LAST_ROUNDS, "phase1 visits/sets/changed", list(st8)[:3], "phase2", list(st8)[4:7], "tiles", (n // 16) ** 3)
