"""Shared driver for the world_size-2 protocol tests (gloo): every rank builds its extended
slabs from one seeded global volume, runs the sharded op, and rank 0 compares the
concatenated result with the oracle on the whole volume."""
import os
import sys
import traceback
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def global_volume(shape=(23, 20, 45), seed=0):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    f = ndimage.gaussian_filter(rng.normal(size=shape), 1.5)
    return (f / np.abs(f).max() * 2000 + 300).astype(np.int16)


def ext_slab(a: np.ndarray, shard, lo=True, hi=True):
    z0 = shard.z0 - (1 if (lo and shard.has_lo) else 0)
    z1 = shard.z1 + (1 if (hi and shard.has_hi) else 0)
    return np.ascontiguousarray(a[z0:z1])


def _worker(rank, world, port, fn_name, module, device, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        if device == "nccl":
            torch.cuda.set_device(rank)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        else:
            if device == "cuda":
                torch.cuda.set_device(rank % torch.cuda.device_count())
            dist.init_process_group("gloo", rank=rank, world_size=world)
        mod = __import__(module)
        res = getattr(mod, fn_name)(rank, world, device)
        q.put((rank, "ok", res))
    except Exception:
        q.put((rank, "error", traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def run_ranks(fn_name, module, world=2, device="cpu", timeout=240):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, module, device, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in range(world):
            rank, status, res = q.get(timeout=timeout)
            assert status == "ok", f"rank {rank} failed:\\n{res}"
            out[rank] = res
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    return out
