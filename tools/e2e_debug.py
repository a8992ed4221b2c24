import sys, time, os
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scipy.ndimage import generate_binary_structure
from invesalius3_b200 import device as dev, phantom, invesalius_rs
import bench
n = 512
print(bench.bind_to_gpu_numa(0) if "--bind" in sys.argv else "unbound", "threads", torch.get_num_threads())
vol, seeds = bench.make_volume(n); seed = seeds[0]
st = generate_binary_structure(3, 1)
h_ext = torch.from_numpy(vol).pin_memory()
h_out = torch.zeros((n, n, n), dtype=torch.uint8).pin_memory()
np_vol, np_out = h_ext.numpy(), h_out.numpy()
def T(f, reps=5):
    f(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        t0=time.perf_counter(); r=f(); torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
    return f"min {min(ts):7.2f} mean {sum(ts)/len(ts):7.2f} ms"
print("zero_          ", T(lambda: h_out.zero_()))
print("to_device(vol) ", T(lambda: dev.to_device(np_vol)))
print("to_device(out) ", T(lambda: dev.to_device(np_out)))
d = dev.to_device(np_vol); o = dev.to_device(np_out)
print("flood (device) ", T(lambda: (o.zero_(), dev.floodfill_threshold(d, [seed], 226, 3071, 254, st, o))))
print("to_host(out)   ", T(lambda: dev.to_host(o, np_out)))
print("shares_memory  ", T(lambda: np.shares_memory(np_vol, np_out)))
print("shim call      ", T(lambda: (h_out.zero_(), invesalius_rs.floodfill_threshold(np_vol, [seed], 226, 3071, 254, st, np_out))))
def seq():
    d = dev.to_device(np_vol); o = dev.to_device(np_out)
    dev.floodfill_threshold(d, [seed], 226, 3071, 254, st, o)
    dev.to_host(o, np_out)
print("manual sequence", T(seq))
def seq2():
    d = dev.to_device(np_vol); torch.cuda.synchronize(); t1 = time.perf_counter()
    o = dev.to_device(np_out); torch.cuda.synchronize(); t2 = time.perf_counter()
    dev.floodfill_threshold(d, [seed], 226, 3071, 254, st, o); t3 = time.perf_counter()
    dev.to_host(o, np_out); t4 = time.perf_counter()
    return t2 - t1, t3 - t2, t4 - t3
seq2(); r = [seq2() for _ in range(5)]
print("in-sequence: to_device(out) %.2f flood %.2f to_host %.2f ms" % tuple(1e3 * min(x[i] for x in r) for i in range(3)))
