import sys
from pathlib import Path; sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from scipy import ndimage
from scipy.ndimage import generate_binary_structure
from invesalius3_b200 import labeling, phantom
for nzs in (8, 40, 41, 96):
    vol = phantom.ct((nzs, 512, 512), seed=2)
    img = (vol >= 226) & (vol <= 3071)
    st = generate_binary_structure(3, 1)
    want, n = ndimage.label(img, st, output=np.uint32)
    got, m = labeling.label(img, st)
    eq = np.array_equal(got, want)
    print(nzs, "n", n, m, "equal", eq, "voxels", img.size, "blocks", -(-img.size // 2048))
    if not eq:
        bad = np.flatnonzero(got.ravel() != want.ravel())
        print("  first bad flat index", bad[0], "got", got.ravel()[bad[0]], "want", want.ravel()[bad[0]], "count", bad.size)
        pairs = np.unique(np.stack([got.ravel()[img.ravel()], want.ravel()[img.ravel()]]), axis=1)
        print("  distinct (got, want) pairs", pairs.shape[1], "bijective", pairs.shape[1] == n)
        # first label where numbering diverges
        first_of_want = np.zeros(n + 1, np.int64); first_of_got = np.zeros(m + 1, np.int64)
        idx = np.flatnonzero(img.ravel())
        fw = np.full(n + 1, img.size, np.int64); np.minimum.at(fw, want.ravel()[idx], idx)
        fg = np.full(m + 1, img.size, np.int64); np.minimum.at(fg, got.ravel()[idx], idx)
        print("  want first-voxel order monotone", np.all(np.diff(fw[1:]) > 0), "got monotone", np.all(np.diff(fg[1:]) > 0))
        d = np.flatnonzero(fw[1:] != fg[1:])
        print("  first diverging label", d[:5] + 1, fw[1:][d[:5]], fg[1:][d[:5]])
import time, torch
from invesalius3_b200 import device as dev
vol = phantom.ct((512, 512, 512), seed=2)
for name, img in (("bone threshold mask", (vol >= 226) & (vol <= 3071)), ("its complement (fill-holes input)", ~((vol >= 226) & (vol <= 3071)))):
    fgt = torch.from_numpy(np.ascontiguousarray(img).view(np.uint8)).cuda()
    labeling.label_device(fgt, st)
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lab, nn = labeling.label_device(fgt, st); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t0 = time.perf_counter(); w, nw = ndimage.label(img, st, output=np.uint32); cpu = time.perf_counter() - t0
    print(f"512^3 {name}: device {min(ts):.2f} ms ({img.size / min(ts) / 1e3:.0f} Mvoxel/s), scipy {cpu * 1e3:.0f} ms, labels {nn} == {nw}, equal {np.array_equal(lab.cpu().numpy().view(np.uint32), w)}")
