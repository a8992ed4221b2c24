#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json configs[1]):
threshold -> seeded flood-fill region grow -> marching cubes on a 512^3 int16 CT phantom.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One JSON line on stdout (rank 0). `value` = Mvoxel/s of the whole job with inputs resident in
HBM (CUDA-event timed, max over ranks); `e2e` = the same three ops called through the
reference-shaped numpy API (slice_ops / invesalius_rs / surface_process) on pinned HOST
buffers, host<->device copies inside the timed region; `roofline` = the dominant kernel
against the measured HBM peak; `cpu_baseline` = the CPU restatement of the reference path
timed on this box's host cores on a bounded sample (a reported baseline, not the target).
`--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Mvoxels/s threshold+floodfill+MC on 512^3 int16"
UNIT = "Mvoxel/s"
THR = (226, 3071)          # presets.py "Bone"
SPACING = (1.0, 1.0, 1.0)
SEED_SLICE_FRAC = 0.5
FILL = 254


def workload_desc(n):
    return (f"{n}^3 synthetic int16 CT phantom (seed 2) per GPU: threshold [226,3071] -> 6-connected flood fill "
            f"from one seed per {n}^3 shard -> marching cubes iso 127 on the grown mask")


def measured_peak():
    try:
        return float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def make_volume(n, world=1, zrange=None):
    """The job's volume is (n*world) x n x n, Z-sharded; returns the planes `zrange` (default:
    all) and the global seeds (x, y, z): one per n^3 shard, in the shard's middle slice (the
    first in-range voxel in raveled order). One connected region spans all shards, so the
    waves started in different shards meet at the shard boundaries."""
    from invesalius3_b200 import phantom
    DZ = n * world
    seeds = []
    for r in range(world):
        zmid = r * n + int(n * SEED_SLICE_FRAC)
        mid = phantom.ct((DZ, n, n), seed=2, zrange=(zmid, zmid + 1))
        sx, sy, _ = phantom.first_seed_in_range(mid, 0, *THR)
        seeds.append((sx, sy, zmid))
    vol = phantom.ct((DZ, n, n), seed=2, zrange=zrange)
    return vol, seeds


# ------------------------------------------------------------------ CPU reference path
def cpu_step(vol, seed, threads):
    """The reference's CPU path restated (oracle/): NumPy threshold statements verbatim
    (single thread, as in the reference), serial stack flood fill, marching cubes over
    20(+1)-slice Z pieces on a thread pool (surface.py:1360-1381 uses a process pool)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from scipy.ndimage import generate_binary_structure
    dz, dy, dx = vol.shape
    mm = np.zeros((dz + 1, dy + 1, dx + 1), np.uint8)
    oracle.set_mask_threshold_numpy(vol, mm, THR)
    out = np.zeros(vol.shape, np.uint8)
    oracle.floodfill_threshold(vol, [seed], THR[0], THR[1], FILL, generate_binary_structure(3, 1), out)
    n_pieces = int(round(dz / 20 + 0.5))
    rois = [slice(i * 20, min(dz, (i + 1) * 20 + 1)) for i in range(n_pieces) if i * 20 < dz]

    def piece(roi):
        a = np.ascontiguousarray(out[roi])
        if a.shape[0] < 2:
            return 0
        v, f = oracle.marching_cubes(a, 127, SPACING, (0, 0, roi.start), True)
        return len(f)

    with ThreadPoolExecutor(max_workers=max(1, min(threads, len(rois)))) as ex:
        ntri = sum(ex.map(piece, rois))
    return ntri


def cpu_sample(vol, seed, nslices):
    """A slab of the workload around the seed slice (bounded CPU time)."""
    dz = vol.shape[0]
    z0 = max(0, min(dz - nslices, seed[2] - nslices // 2))
    return np.ascontiguousarray(vol[z0:z0 + nslices]), (seed[0], seed[1], seed[2] - z0)


def time_cpu(vol, seed, nslices, reps, threads):
    sub, sseed = cpu_sample(vol, seed, nslices)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_step(sub, sseed, threads)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    return sub.size / t / 1e6, f"{sub.shape[0]}x{sub.shape[1]}x{sub.shape[2]} slab around the seed, best of {reps}"


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if "Active" in v and "Not" not in v:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows)}


def bind_to_gpu_numa(local_rank):
    """Best effort: run this process on the CPUs of the NUMA node the GPU hangs off, so that
    first-touch puts the pinned host buffers next to the GPU's PCIe root (cross-socket DMA and
    memset are several times slower). Returns a short description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return f"numa: bound to node {node} ({len(ids)} cpus)"
    except Exception as e:   # noqa: BLE001
        return f"numa: not bound ({type(e).__name__})"


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from scipy.ndimage import generate_binary_structure
    from invesalius3_b200 import _lib, device as dev, invesalius_rs, slice_ops, surface_process
    from invesalius3_b200.mesh import marching_cubes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev.require_cuda()
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    torch.set_num_threads(16)   # host memset of the out mask: 128 OpenMP threads make it erratic (0.3 .. 25 ms)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    n = args.size
    from invesalius3_b200 import dist as zd
    shard = zd.ZShard(n * world, rank, world)
    ext_np, seeds = make_volume(n, world, (shard.ze0, shard.ze1))   # own planes + halo planes
    seed = seeds[0]
    vol = shard.interior(torch.from_numpy(ext_np)).numpy()          # this rank's n^3 shard
    strct = generate_binary_structure(3, 1)
    N = vol.size
    nz_ext = ext_np.shape[0]

    # pinned host buffers for the e2e leg (numpy views of torch pinned tensors)
    h_ext = torch.from_numpy(ext_np).pin_memory()
    h_vol = shard.interior(h_ext)
    h_mask = torch.zeros((n + 1, n + 1, n + 1), dtype=torch.uint8).pin_memory()
    h_out = torch.zeros((nz_ext, n, n), dtype=torch.uint8).pin_memory()
    np_vol, np_mask, np_out = h_vol.numpy(), h_mask.numpy(), h_out.numpy()

    d_ext = h_ext.cuda(non_blocking=True)
    d_vol = shard.interior(d_ext)
    d_mask = torch.empty((n, n, n), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((nz_ext, n, n), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    info = {}

    def do_flood(data_ext, out_ext):
        if world == 1:
            info["rounds"] = dev.floodfill_threshold(data_ext, [seed], THR[0], THR[1], FILL, strct, out_ext)
        else:
            info["rounds"] = zd.floodfill_threshold(data_ext, seeds, THR[0], THR[1], FILL, strct, out_ext, shard)

    def do_surface(out_ext):
        if world == 1:
            v, f = marching_cubes(out_ext, 127, SPACING, (0, 0, 0), True)
            info["V"], info["T"] = int(v.shape[0]), int(f.shape[0])
        else:
            v, f, _, info["V"], info["T"] = zd.marching_cubes(out_ext[int(shard.has_lo):], 127, SPACING, (0, 0, 0),
                                                              True, shard)
        return v, f

    def flood_and_surface(data_ext, out_ext):
        do_flood(data_ext, out_ext)
        return do_surface(out_ext)

    def step_device(ev=None):
        if ev: ev[0].record()
        dev.threshold(d_vol, THR[0], THR[1], out=d_mask)
        if ev: ev[1].record()
        d_out.zero_()
        do_flood(d_ext, d_out)
        if ev: ev[2].record()
        do_surface(d_out)
        if ev: ev[3].record()

    e2e_min = {}
    e2e_calls = {"set_mask_threshold": 0.0, "zero_out_mask": 0.0, "floodfill_threshold": 0.0, "contour": 0.0}

    def step_e2e():
        if world == 1:
            t0 = time.perf_counter()
            slice_ops.set_mask_threshold(np_vol, np_mask, THR)
            t1 = time.perf_counter()
            h_out.zero_()      # the reference allocates out_mask = np.zeros_like(mask) here (styles.py:3183)
            t2 = time.perf_counter()
            invesalius_rs.floodfill_threshold(np_vol, [seed], THR[0], THR[1], FILL, strct, np_out)
            t3 = time.perf_counter()
            v, f = surface_process.contour(np_out, [127], SPACING, 0, True)
            t4 = time.perf_counter()
            for k, dt in zip(e2e_calls, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                e2e_calls[k] += dt
                e2e_min[k] = min(e2e_min.get(k, 1e9), dt)
            return v, f
        # N > 1: the sharded pipeline fed from / drained to pinned host memory
        t_ext = dev.to_device(h_ext.numpy())
        m = dev.threshold(shard.interior(t_ext), THR[0], THR[1])
        dev.to_host(m, np_mask[1:, 1:, 1:])
        np_mask[1:, 0, 0] = 1
        o_ext = torch.zeros((nz_ext, n, n), dtype=torch.uint8, device="cuda")
        v, f = flood_and_surface(t_ext, o_ext)
        dev.to_host(shard.interior(o_ext), shard.interior(h_out).numpy())
        return v.cpu().numpy(), f.cpu().numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    lib.b2v_launch_count_reset()
    stage_ms = np.zeros(3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device(ev)
        torch.cuda.synchronize()
        stage_ms += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    e1.record()
    barrier()
    launches = int(lib.b2v_launch_count())
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / args.steps
    value = world * N / (ms_per_step * 1e-3) / 1e6
    stage_ms /= args.steps

    # ---- e2e leg: reference-shaped numpy API on pinned host buffers
    e2e_steps = max(1, min(args.steps, 5))
    v, f = step_e2e()
    v, f = step_e2e()   # results stay bound across calls, as in the timed loop: the pinned result
    v, f = step_e2e()   # pool reaches its steady state (two sets of blocks in flight)
    for k in e2e_calls:
        e2e_calls[k] = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        v, f = step_e2e()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    if world == 1:
        h2d = 2 * N * 2 + 2 * N            # image twice (threshold, flood fill), out in, out again for MC
    else:
        h2d = ext_np.nbytes                # the extended slab once; the pipeline stays on the device
    d2h = 2 * N + v.nbytes + f.nbytes  # mask, out, mesh
    e2e_value = world * N / e2e_s / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = measured_peak()
    # dominant stage and its roofline (algorithmic bytes: SURVEY.md 8d / DESIGN.md)
    alg = {"threshold": 3.0 * N, "floodfill": 4.0 * N,
           "marching_cubes": 1.0 * N + (12.0 * info["V"] + 12.0 * info["T"]) / world}
    names = list(alg)
    dom = int(np.argmax(stage_ms))
    achieved = alg[names[dom]] / (stage_ms[dom] * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes per launch of that stage from the committed ncu capture (profiles/)
        traffic = json.load(open(ROOT / "profiles" / "r01_traffic.json"))["stages"][names[dom]]["traffic"]
    except Exception:
        pass
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cpu_baseline = None
    if world == 1:   # the CPU baseline is an N=1 figure (rank 0's host cores)
        cpu_v, cpu_sample_desc = time_cpu(vol, seed, min(n, args.cpu_slices), 1, cores)
        cpu_baseline = {"value": round(cpu_v, 2), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": cpu_sample_desc}
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": workload_desc(n),
                   "volume": f"{n * world}x{n}x{n} (Z-sharded, one halo plane per inner side)",
                   "shard": f"{n}^3 voxels per GPU", "l2": "inputs (256 MiB int16 + 128 MiB uint8) exceed the 126 MB L2",
                   "host": numa, "flood_rounds": info["rounds"], "vertices": info["V"], "triangles": info["T"],
                   "stage_ms": {k: round(float(m), 4) for k, m in zip(names, stage_ms)}},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": round(e2e_s * 1e3, 3),
                "ms_per_call": {k: round(v / e2e_steps * 1e3, 3) for k, v in e2e_calls.items()},
                "ms_per_call_min": {k: round(v * 1e3, 3) for k, v in e2e_min.items()},
                "api": ("slice_ops.set_mask_threshold + invesalius_rs.floodfill_threshold + surface_process.contour "
                        "on pinned numpy buffers") if world == 1 else
                       "dist.* sharded pipeline fed from / drained to pinned host buffers"},
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": round(achieved, 1), "peak": peak,
                     "peak_kind": peak_kind, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                     "note": "the dominant stage is the flood fill: 4 B/voxel over the whole op (build + rounds + "
                             "write); its rounds are L2/latency-bound, see DESIGN.md",
                     "per_stage_GBs": {k: round(alg[k] / (m * 1e-3) / 1e9, 1) for k, m in zip(names, stage_ms)}},
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.size
    vol, seeds = make_volume(n)
    seed = seeds[0]
    cores = os.cpu_count() or 1
    sub, sseed = cpu_sample(vol, seed, min(n, args.cpu_slices))
    for _ in range(min(args.warmup, 1)):
        cpu_step(sub, sseed, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(sub, sseed, cores)
    s = (time.perf_counter() - t0) / args.steps
    v = sub.size / s / 1e6
    sample = f"{sub.shape[0]}x{sub.shape[1]}x{sub.shape[2]} slab of the {n}^3 workload around the seed, per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": round(s * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": workload_desc(n)},
        "cpu_baseline": {"value": round(v, 2), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--cpu-slices", type=int, default=128)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
