"""Persistent 6-connected watershed engine against the generic round kernels (same fixpoint: labels and
ambiguous masks must be identical), both algorithms, a few shapes; then 512^3 timings."""
import os, sys, subprocess, json
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scipy.ndimage import generate_binary_structure
from invesalius3_b200 import phantom, watershed_process as wp
sys.path.insert(0, str(Path(__file__).resolve().parent))
from ws_bench import markers_for

st = generate_binary_structure(3, 1)

def run(vol, mk, alg, ww_wl=True):
    r = wp.watershed_device(torch.from_numpy(vol).cuda(), torch.from_numpy(mk).cuda(), st, alg, 3, ww_wl, -18, 406,
                            return_ambiguous=True)
    return r[0].cpu().numpy(), r[1].cpu().numpy()

if len(sys.argv) > 1 and sys.argv[1] == "child":
    shapes = json.loads(sys.argv[2])
    out = {}
    for sh in shapes:
        vol = phantom.ct(tuple(sh), seed=4)
        mk = markers_for(vol, 4)
        for alg in ("Watershed", "Watershed IFT"):
            for wwwl in (True, False):
                l, a = run(vol, mk, alg, wwwl)
                tag = os.environ.get("B2V_WS_GENERIC", "0")
                np.save(f"/tmp/ws_{tag}_{'x'.join(map(str, sh))}_{alg.replace(' ', '')}_{int(wwwl)}.npy", np.stack([l.astype(np.int16), a.astype(np.int16)]))
    sys.exit(0)

shapes = [[40, 50, 70], [33, 65, 129], [96, 96, 96], [1, 80, 90]]
for env in ("0", "1"):
    e = dict(os.environ)
    if env == "1":
        e["B2V_WS_GENERIC"] = "1"
    else:
        e.pop("B2V_WS_GENERIC", None)
        e["B2V_WS_GENERIC_TAG"] = "0"
    subprocess.run([sys.executable, __file__, "child", json.dumps(shapes)], check=True, env=e)
ok = True
for sh in shapes:
    for alg in ("Watershed", "WatershedIFT"):
        for wwwl in (1, 0):
            a = np.load(f"/tmp/ws_0_{'x'.join(map(str,sh))}_{alg}_{wwwl}.npy")
            b = np.load(f"/tmp/ws_1_{'x'.join(map(str,sh))}_{alg}_{wwwl}.npy")
            same_l, same_a = np.array_equal(a[0], b[0]), np.array_equal(a[1], b[1])
            ok &= same_l and same_a
            print(sh, alg, wwwl, "labels equal", same_l, "ambiguous equal", same_a, "diff", int((a[0] != b[0]).sum()), int((a[1] != b[1]).sum()))
print("ALL EQUAL" if ok else "MISMATCH")
n = 512
vol = phantom.ct((n, n, n), seed=4)
mk = markers_for(vol, 4)
t_vol, t_mk = torch.from_numpy(vol).cuda(), torch.from_numpy(mk).cuda()
for alg in ("Watershed", "Watershed IFT"):
    for amb in (False, True):
        wp.watershed_device(t_vol, t_mk, st, alg, 3, True, -18, 406, return_ambiguous=amb)
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); wp.watershed_device(t_vol, t_mk, st, alg, 3, True, -18, 406, return_ambiguous=amb); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(alg, "ambiguous" if amb else "labels only", f"{min(ts):.2f} ms", "rounds", wp.LAST_ROUNDS if hasattr(wp, "LAST_ROUNDS") else "")
