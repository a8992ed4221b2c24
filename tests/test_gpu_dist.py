"""Z-sharded ops with the CUDA backend: 2 ranks. On a 1-GPU box both ranks share the device
and talk through gloo (tensors staged via the host); with >= 2 GPUs the same functions also
run over NCCL, one rank per GPU. Results must equal the oracle on the whole volume."""
import numpy as np
import pytest
import torch

from dist_common import ext_slab, global_volume, run_ranks
from test_dist_gloo import THR, ff_cases

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def rank_all(rank, world, device):
    from invesalius3_b200 import dist as d
    g = global_volume((37, 40, 96), seed=3)
    shard = d.ZShard(g.shape[0], rank, world)
    res = {}
    # NVLink peer mailboxes (one rank per GPU, NCCL job): the fused exchange paths are run as well
    link = d.peer_link(shard, g.shape[1], g.shape[2]) if device == "nccl" else None
    res["link"] = link is not None
    own = torch.from_numpy(np.ascontiguousarray(g[shard.z0:shard.z1])).to(_dev())
    for axis in (0, 1, 2):
        for kind in ("max", "min", "mean"):
            res[("mip", axis, kind)] = d.mip(own, axis, kind, shard).cpu().numpy()
    res["thr"] = (shard.z0, shard.z1, d.threshold(own, *THR, shard).cpu().numpy())
    for axis in (0, 1, 2):   # axis 0: the rays cross the shards (state hand-off)
        res[("mida", axis)] = d.mida(own, axis, 300, 600, shard).cpu().numpy()
        res[("lmip", axis)] = d.lmip(own, axis, 700, 3033, shard).cpu().numpy()
    # contour-MIP on the extended slab (n = 1: the power is exact, so the result is bit-exact)
    ext = torch.from_numpy(ext_slab(g, shard).copy()).to(_dev())
    for axis in (0, 1, 2):   # axis 0: rays cross the shards; tmip 2: extrema of the contour volume all-reduced
        for tmip in (0, 1, 2):
            res[("fcm", axis, tmip)] = d.fast_countour_mip(ext, 1.0, axis, 300, 600, tmip, shard).cpu().numpy()
    # fill holes: labels of the whole mask, sizes summed over the shards
    from scipy import ndimage
    hm = (g > 600).astype(np.uint8) * 255
    lab, nlab = ndimage.label(hm == 0, ndimage.generate_binary_structure(3, 1), output=np.uint32)
    for max_size in (3, 50):
        m = torch.from_numpy(hm[shard.z0:shard.z1].copy()).to(_dev())
        lt = torch.from_numpy(lab[shard.z0:shard.z1].view(np.int32).copy()).to(_dev())
        ret = d.fill_holes_automatically(m, lt, int(nlab), max_size, shard)
        res[("fh", max_size)] = (ret, m.cpu().numpy())
    # flood fill
    for ci, (strct, seeds) in enumerate(ff_cases(g)):
        data = torch.from_numpy(ext_slab(g, shard)).to(_dev())
        dext = torch.empty_like(data)
        d.shard_copy = None
        out_g = np.zeros(g.shape, np.uint8); out_g[:, 10, :] = 254
        out = torch.from_numpy(ext_slab(out_g, shard).copy()).to(_dev())
        # halo planes arrive by exchange, not from the global array
        if shard.has_lo: data[0] = 0; out[0] = 0
        if shard.has_hi: data[-1] = 0; out[-1] = 0
        d.exchange_halo(data, shard); d.exchange_halo(out, shard)
        out2 = out.clone()
        outer = d.floodfill_threshold(data, seeds, 100, 3071, 254, strct, out, shard)
        res[("ff", ci)] = (shard.interior(out).cpu().numpy(), outer)
        if link is not None:
            outer = d.floodfill_threshold(data, seeds, 100, 3071, 254, strct, out2, shard, link=link)
            res[("ffp", ci)] = (shard.interior(out2).cpu().numpy(), outer)
    # marching cubes on the thresholded mask
    mask = ((g >= THR[0]) & (g <= THR[1])).astype(np.uint8) * 255
    vol = torch.from_numpy(ext_slab(mask, shard, lo=False, hi=True)).to(_dev())
    v, t, vbase, tv, tt = d.marching_cubes(vol, 127, (0.5, 0.75, 1.5), (-1, -1, 3), True, shard)
    res["mc"] = (v.cpu().numpy(), t.cpu().numpy(), vbase, tv, tt)
    if link is not None:
        for rep in range(3):    # epochs alternate the mailbox halves
            v, t, vbase, tv, tt = d.marching_cubes(vol, 127, (0.5, 0.75, 1.5), (-1, -1, 3), True, shard, link=link)
        res["mcp"] = (v.cpu().numpy(), t.cpu().numpy(), vbase, tv, tt)
        link.close()
    return res


def _check(out, orc):
    g = global_volume((37, 40, 96), seed=3)
    for rank in (0, 1):
        for axis in (0, 1, 2):
            for kind in ("max", "min", "mean"):
                want = {"max": g.max, "min": g.min, "mean": g.mean}[kind](axis)
                assert np.array_equal(out[rank][("mip", axis, kind)], want), (rank, axis, kind)
        for axis in (0, 1, 2):
            want = np.zeros([(40, 96), (37, 96), (37, 40)][axis], np.int16)
            orc.mida(g, axis, 300, 600, want)
            assert np.array_equal(out[rank][("mida", axis)], want), ("mida", rank, axis)
            orc.lmip(g, axis, 700, 3033, want)
            assert np.array_equal(out[rank][("lmip", axis)], want), ("lmip", rank, axis)
            for tmip in (0, 1, 2):
                orc.fast_countour_mip(g, 1.0, axis, 300, 600, tmip, want)
                assert np.array_equal(out[rank][("fcm", axis, tmip)], want), ("fcm", rank, axis, tmip)
        z0, z1, m = out[rank]["thr"]
        want = np.zeros(g.shape, np.uint8)
        orc.threshold(g, *THR, want, False)
        assert np.array_equal(m, want[z0:z1])
    from scipy import ndimage
    hm = (g > 600).astype(np.uint8) * 255
    lab, nlab = ndimage.label(hm == 0, ndimage.generate_binary_structure(3, 1), output=np.uint32)
    for max_size in (3, 50):
        want = hm.copy()
        ret = orc.fill_holes_automatically(want, lab, int(nlab), max_size)
        assert np.array_equal(np.concatenate([out[r][("fh", max_size)][1] for r in (0, 1)]), want), max_size
        assert out[0][("fh", max_size)][0] == out[1][("fh", max_size)][0] == ret
    for ci, (strct, seeds) in enumerate(ff_cases(g)):
        want = np.zeros(g.shape, np.uint8); want[:, 10, :] = 254
        orc.floodfill_threshold(g, seeds, 100, 3071, 254, strct, want)
        got = np.concatenate([out[r][("ff", ci)][0] for r in (0, 1)])
        assert np.array_equal(got, want), ci
        if out[0]["link"]:
            got = np.concatenate([out[r][("ffp", ci)][0] for r in (0, 1)])
            assert np.array_equal(got, want), ("peer", ci)
    mask = ((g >= THR[0]) & (g <= THR[1])).astype(np.uint8) * 255
    V, T = orc.marching_cubes(mask, 127, (0.5, 0.75, 1.5), (-1, -1, 3), True)
    gv = np.concatenate([out[0]["mc"][0], out[1]["mc"][0]])
    gt = np.concatenate([out[0]["mc"][1], out[1]["mc"][1]])
    assert out[0]["mc"][3] == len(V) and out[0]["mc"][4] == len(T)
    assert np.array_equal(gv, V) and np.array_equal(gt.astype(np.int64), T)
    if out[0]["link"]:
        gv = np.concatenate([out[0]["mcp"][0], out[1]["mcp"][0]])
        gt = np.concatenate([out[0]["mcp"][1], out[1]["mcp"][1]])
        assert out[1]["mcp"][3] == len(V) and out[1]["mcp"][4] == len(T)
        assert np.array_equal(gv, V) and np.array_equal(gt.astype(np.int64), T)


def test_sharded_ops_two_ranks_one_gpu_gloo(orc):
    _check(run_ranks("rank_all", "test_gpu_dist", device="cuda"), orc)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_ops_two_ranks_nccl(orc):
    out = run_ranks("rank_all", "test_gpu_dist", device="nccl")
    assert out[0]["link"] and out[1]["link"], "peer mailboxes (cudaIpc over NVLink) could not be set up"
    _check(out, orc)


# ---- watershed over Z shards (BASELINE configs[3]: boundary planes exchanged until nothing improves)
def rank_watershed(rank, world, device):
    from scipy.ndimage import generate_binary_structure
    from invesalius3_b200 import dist as d, phantom
    g = phantom.ct((40, 48, 64), seed=4)
    mk = np.zeros(g.shape, np.uint8)
    mk[20, 24, 30:34] = 1; mk[2, 2, 2:6] = 2; mk[30, 40, 50:54] = 1; mk[38, 5, 60:63] = 2
    shard = d.ZShard(g.shape[0], rank, world)
    st6 = generate_binary_structure(3, 1)
    res = {}
    for alg in ("Watershed", "Watershed IFT"):
        for ww_wl in (True, False):
            img = torch.from_numpy(ext_slab(g, shard)).to(_dev())
            m = torch.from_numpy(ext_slab(mk, shard)).to(_dev())
            lab, amb, ex = d.watershed(img, m, st6, alg, 3, ww_wl, -18, 406, shard, return_ambiguous=True)
            res[(alg, ww_wl)] = (lab.cpu().numpy(), amb.cpu().numpy(), ex)
    return res


def _check_watershed(out, world):
    from scipy.ndimage import generate_binary_structure
    from invesalius3_b200 import phantom, watershed_process as wp
    g = phantom.ct((40, 48, 64), seed=4)
    mk = np.zeros(g.shape, np.uint8)
    mk[20, 24, 30:34] = 1; mk[2, 2, 2:6] = 2; mk[30, 40, 50:54] = 1; mk[38, 5, 60:63] = 2
    st6 = generate_binary_structure(3, 1)
    for alg in ("Watershed", "Watershed IFT"):
        for ww_wl in (True, False):
            want = wp.watershed_device(torch.from_numpy(g).cuda(), torch.from_numpy(mk).cuda(), st6, alg, 3, ww_wl, -18, 406,
                                       return_ambiguous=True)
            lab = np.concatenate([out[r][(alg, ww_wl)][0] for r in range(world)])
            amb = np.concatenate([out[r][(alg, ww_wl)][1] for r in range(world)])
            assert np.array_equal(lab, want[0].cpu().numpy()), (alg, ww_wl)
            assert np.array_equal(amb, want[1].cpu().numpy()), (alg, ww_wl)
            assert out[0][(alg, ww_wl)][2] >= 3


def test_sharded_watershed_two_ranks_one_gpu_gloo():
    _check_watershed(run_ranks("rank_watershed", "test_gpu_dist", device="cuda"), 2)


def test_sharded_watershed_three_ranks_one_gpu_gloo():
    _check_watershed(run_ranks("rank_watershed", "test_gpu_dist", world=3, device="cuda"), 3)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_watershed_two_ranks_nccl():
    _check_watershed(run_ranks("rank_watershed", "test_gpu_dist", device="nccl"), 2)
