"""BASELINE configs [2], [3], [4] on N GPUs of one box (torchrun, one rank per GPU), verified against a
single-GPU run of the gathered volume on rank 0:

  C3  1024^3 int16: MaxIP / MinIP / MeanIP and MIDA on the three axes, Z-sharded (dist.mip, dist.mida)
  C4  512^3 watershed, 8 marker balls, ww 406 / wl -18, both algorithms (dist.watershed)
  C5  2048 x 2048 x 1024 int16: threshold -> marching cubes with the boundary stitch (dist.threshold,
      dist.marching_cubes over the peer mailboxes)

    python -m torch.distributed.run --nproc-per-node N tools/configs_multigpu.py [--scale 1.0] [--skip c5]

Device-timed (CUDA events, barrier + synchronize on both sides, max over ranks, best of 3). The
volumes are device-generated analytic phantoms (shell + spheres + texture + hash noise), identical on
every rank for the planes it owns. One JSON object on stdout (rank 0)."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scipy.ndimage import generate_binary_structure  # noqa: E402
from invesalius3_b200 import device as dev, dist as zd, projection, watershed_process as wp  # noqa: E402
from invesalius3_b200.mesh import marching_cubes  # noqa: E402


def phantom_planes(DZ, dy, dx, z0, z1, seed):
    """int16 [z1 - z0][dy][dx] on the device: the planes z0 .. z1 of a DZ-plane analytic head phantom."""
    z = torch.arange(z0, z1, dtype=torch.float32, device="cuda")[:, None, None]
    y = torch.arange(dy, dtype=torch.float32, device="cuda")[None, :, None]
    x = torch.arange(dx, dtype=torch.float32, device="cuda")[None, None, :]
    r2 = ((z - (DZ - 1) / 2) / DZ) ** 2 + ((y - (dy - 1) / 2) / dy) ** 2 + ((x - (dx - 1) / 2) / dx) ** 2
    v = torch.full(r2.shape, -1000.0, device="cuda")
    v = torch.where(r2 <= 0.42 ** 2, torch.full_like(v, 40.0), v)
    v = torch.where((r2 <= 0.40 ** 2) & (r2 >= 0.36 ** 2), torch.full_like(v, 1200.0), v)
    inner = r2 < 0.36 ** 2
    v = torch.where(inner, v + 300.0 * torch.sin(0.35 * x) * torch.sin(0.31 * y) * torch.sin(0.27 * z), v)
    g = torch.Generator().manual_seed(seed)
    for c in torch.rand((8, 3), generator=g).mul(0.4).add(0.3).tolist():
        d2 = (z - c[0] * DZ) ** 2 + (y - c[1] * dy) ** 2 + (x - c[2] * dx) ** 2
        v = torch.where((d2 <= (0.06 * min(DZ, dy, dx)) ** 2) & inner, torch.full_like(v, 700.0), v)
    # hash noise: a function of the global voxel index only
    idx = (z.to(torch.int64) * dy + y.to(torch.int64)) * dx + x.to(torch.int64)
    h = (idx * 2654435761 + seed * 40503) & 0xFFFFFFFF
    h = ((h ^ (h >> 15)) * 2246822519) & 0xFFFFFFFF
    h = (h ^ (h >> 13)) & 0xFFFF
    v = v + (h.to(torch.float32) / 65535.0 - 0.5) * 80.0
    return v.clamp_(-1024, 3071).round_().to(torch.int16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every dimension (smoke runs)")
    ap.add_argument("--skip", default="", help="comma list of c3,c4,c5")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    real_stdout = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
    skip = set(args.skip.split(","))
    res = {"n_gpus": world, "scale": args.scale}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=3):
        fn()
        best = 1e30
        for _ in range(reps):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = fn(); b.record(); barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best, out

    def gather_volume(own, DZ):
        """whole [DZ][dy][dx] volume on rank 0 (None elsewhere)."""
        if world == 1:
            return own
        shard = zd.ZShard(DZ, rank, world)
        if rank == 0:
            whole = torch.empty((DZ,) + tuple(own.shape[1:]), dtype=own.dtype, device="cuda")
            whole[: own.shape[0]].copy_(own)
            for r in range(1, world):
                a, b = shard.bounds(r)
                dist.recv(whole[a:b].view(torch.uint8), src=r)
            return whole
        dist.send(own.contiguous().view(torch.uint8), dst=0)
        return None

    # ---------------- C3: projections of 1024^3
    if "c3" not in skip:
        n = max(64, int(1024 * args.scale) // (8 * world) * (8 * world))
        shard = zd.ZShard(n, rank, world)
        own = phantom_planes(n, n, n, shard.z0, shard.z1, 3)
        c3 = {"shape": [n, n, n]}
        outs = {}
        for axis in (0, 1, 2):
            for kind in ("max", "min", "mean"):
                ms, o = timed(lambda: zd.mip(own, axis, kind, shard))
                c3[f"{kind}ip_axis{axis}_ms"] = round(ms, 4); outs[(kind, axis)] = o
            ms, o = timed(lambda: zd.mida(own, axis, 300, 300, shard))
            c3[f"mida_axis{axis}_ms"] = round(ms, 4); outs[("mida", axis)] = o
        N = n ** 3
        c3["maxip_GBs_aggregate"] = [round(2 * N / c3[f"maxip_axis{a}_ms"] / 1e6, 1) for a in (0, 1, 2)]
        whole = gather_volume(own, n)
        if rank == 0:
            ok = True
            for axis in (0, 1, 2):
                for kind in ("max", "min", "mean"):
                    ok &= bool(torch.equal(outs[(kind, axis)], dev.mip(whole, axis, kind)))
                ok &= bool(torch.equal(outs[("mida", axis)], projection.mida(whole, axis, 300, 300)))
            c3["equal_to_single_gpu"] = ok
        del own, whole, outs
        torch.cuda.empty_cache()
        res["c3_projections"] = c3

    # ---------------- C4: watershed 512^3
    if "c4" not in skip:
        n = max(64, int(512 * args.scale) // (16 * world) * (16 * world))
        shard = zd.ZShard(n, rank, world)
        img_ext = phantom_planes(n, n, n, shard.ze0, shard.ze1, 4)
        zz = torch.arange(shard.ze0, shard.ze1, device="cuda")[:, None, None]
        yy = torch.arange(n, device="cuda")[None, :, None]
        xx = torch.arange(n, device="cuda")[None, None, :]
        mk_ext = torch.zeros(img_ext.shape, dtype=torch.uint8, device="cuda")
        g = torch.Generator().manual_seed(44)
        for k in range(8):
            lab = 1 if k < 4 else 2
            r = 0.2 if lab == 1 else 0.47            # foreground balls inside the head, background ones in the air corners
            ang = torch.rand(2, generator=g).mul(6.283).tolist()
            c = [n / 2 + r * n * np.cos(ang[0]) * (1.4 if lab == 2 else 1), n / 2 + r * n * np.sin(ang[0]) * np.cos(ang[1]) * (1.4 if lab == 2 else 1),
                 n / 2 + r * n * np.sin(ang[0]) * np.sin(ang[1]) * (1.4 if lab == 2 else 1)]
            c = [min(max(v, 6.0), n - 7.0) for v in c]
            mk_ext[(zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2 <= 16] = lab
        st6 = generate_binary_structure(3, 1)
        c4 = {"shape": [n, n, n]}
        labs = {}
        for alg in ("Watershed", "Watershed IFT"):
            ms, o = timed(lambda: zd.watershed(img_ext, mk_ext, st6, alg, 3, True, -18, 406, shard), reps=2)
            c4[alg] = {"ms": round(ms, 2), "plane_exchanges": o[1]}
            labs[alg] = o[0]
        whole = gather_volume(shard.interior(img_ext), n)
        mk_whole = gather_volume(shard.interior(mk_ext), n)
        if rank == 0:
            for alg in ("Watershed", "Watershed IFT"):
                t1, want = timed(lambda: wp.watershed_device(whole, mk_whole, st6, alg, 3, True, -18, 406), reps=2) if world == 1 else (None, wp.watershed_device(whole, mk_whole, st6, alg, 3, True, -18, 406))
                c4[alg]["own_planes_equal_single_gpu"] = bool(torch.equal(labs[alg], want[shard.z0:shard.z1]))
        res["c4_watershed"] = c4
        del img_ext, mk_ext, whole, mk_whole, labs
        torch.cuda.empty_cache()

    # ---------------- C5: 2048 x 2048 x 1024 threshold -> marching cubes with the boundary stitch
    if "c5" not in skip:
        dz = max(16 * world, int(1024 * args.scale) // (8 * world) * (8 * world))
        dyx = max(64, int(2048 * args.scale) // 32 * 32)
        shard = zd.ZShard(dz, rank, world)
        img = phantom_planes(dz, dyx, dyx, shard.z0, shard.z1 + int(shard.has_hi), 5)      # own planes + the next shard's first
        link = zd.peer_link(shard, dyx, dyx) if world > 1 else None
        c5 = {"shape": [dz, dyx, dyx], "exchange": link.describe() if link is not None else "none / torch.distributed"}
        mask = torch.empty(img.shape, dtype=torch.uint8, device="cuda")

        def step():
            zd.threshold(img, 226, 3071, shard, out=mask)
            if world == 1:
                v, f = marching_cubes(mask, 127, (1.0, 1.0, 1.0), (0, 0, 0), True)
                return v, f, 0, v.shape[0], f.shape[0]
            return zd.marching_cubes(mask, 127, (1.0, 1.0, 1.0), (0, 0, 0), True, shard, link=link)

        ms, (v, f, vbase, tv, tt) = timed(step)
        nvox = dz * dyx * dyx
        c5.update(ms=round(ms, 3), Mvoxel_s=round(nvox / ms / 1e3, 1), vertices=int(tv), triangles=int(tt))
        tsum = int(f.to(torch.int64).sum().item()); vsum = int(v.contiguous().view(torch.int32).to(torch.int64).sum().item())
        sums = torch.tensor([tsum, vsum, v.shape[0], f.shape[0]], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(sums)
        whole = gather_volume(img[: shard.z1 - shard.z0], dz)
        if rank == 0:
            m1 = dev.threshold(whole, 226, 3071)
            v1, f1 = marching_cubes(m1, 127, (1.0, 1.0, 1.0), (0, 0, 0), True)
            one = [int(f1.to(torch.int64).sum().item()), int(v1.contiguous().view(torch.int32).to(torch.int64).sum().item()), v1.shape[0], f1.shape[0]]
            c5["equal_to_single_gpu"] = one == [int(x) for x in sums.tolist()]
            if world > 1:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); dev.threshold(whole, 226, 3071, out=m1); marching_cubes(m1, 127, (1.0, 1.0, 1.0), (0, 0, 0), True); b.record()
                torch.cuda.synchronize()
                c5["single_gpu_ms_same_volume"] = round(a.elapsed_time(b), 3)
        if link is not None:
            link.close()
        res["c5_threshold_mc"] = c5
    barrier()
    if rank == 0:
        real_stdout.write(json.dumps(res) + "\n"); real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
