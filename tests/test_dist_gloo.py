"""world_size-2 gloo tests (CPU) of the Z-shard protocols in invesalius3_b200/dist.py:
halo exchange, flood-fill plane exchange loop and termination, MIP reductions, and the
marching-cubes numbering / boundary stitch. The compute is delegated to a CPU checker
backend built on the oracle (tests may use the oracle; the product backend is CUDA)."""
import numpy as np
import pytest
import torch

from dist_common import ext_slab, global_volume, run_ranks

THR = (226, 3071)


class CpuBackend:
    """Implements dist.py's backend protocol with NumPy + the oracle."""

    def __init__(self):
        import oracle
        self.o = oracle

    # ---- threshold / projections
    def threshold(self, img, lo, hi, out=None, preserve_markers=False):
        a = img.numpy()
        m = np.zeros(a.shape, np.uint8) if out is None else out.numpy()
        self.o.threshold(a, lo, hi, m, preserve_markers)
        return torch.from_numpy(m)

    def mip(self, img, axis, kind):
        a = img.numpy()
        return torch.from_numpy(np.ascontiguousarray({"max": a.max, "min": a.min, "mean": a.mean}[kind](axis)))

    def sum_axis0(self, img):
        return torch.from_numpy(img.numpy().astype(np.int64).sum(axis=0))

    # ---- rays along z with the state handed from shard to shard (float32 arithmetic in the
    # order of mips.rs:102-168 / :7-86, vectorised over the pixels of a plane)
    def minmax(self, img):
        a = img.numpy()
        return torch.tensor([float(a.min()), float(a.max())], dtype=torch.float32)

    def ray_state(self, img):
        return torch.zeros((3, img.shape[1], img.shape[2]), dtype=torch.int32)

    def mida_z(self, img, wl, ww, minmax, state, first, last):
        a = img.numpy()
        f32 = np.float32
        st = state.numpy().view(np.float32)
        fmax, alpha_p, colour_p = st[0], st[1], st[2]
        mn, mx = f32(minmax[0].item()), f32(minmax[1].item())
        rng = f32(mx - mn)
        inv = f32(1.0) / rng
        wl, ww = f32(a.dtype.type(wl)), f32(a.dtype.type(ww))
        half = ww / f32(2.0)
        wmn, wmx = f32(wl - half), f32(wl + half)
        with np.errstate(all="ignore"):
            for z in range(a.shape[0]):
                act = ~(alpha_p >= f32(1.0))
                vl = a[z].astype(f32)
                fpi = inv * (vl - mn)
                gt = fpi > fmax
                dl = np.where(gt, fpi - fmax, f32(0.0)).astype(f32)
                nfmax = np.where(gt, fpi, fmax).astype(f32)
                bt = f32(1.0) - dl
                alpha = np.where(vl < wmn, f32(0.0), np.where(vl > wmx, f32(1.0), (vl - wmn) / (wmx - wmn))).astype(f32)
                one_m = f32(1.0) - bt * alpha_p
                colour = bt * colour_p + (one_m * fpi) * alpha
                cur = bt * alpha_p + one_m * alpha
                fmax[act] = nfmax[act]
                colour_p[act] = colour[act]
                alpha_p[act] = cur[act]
        if not last:
            return None
        val = rng * colour_p + mn
        odt = np.uint8 if a.dtype == np.float64 else a.dtype
        return torch.from_numpy(np.trunc(val).astype(odt))

    def lmip_z(self, img, tmin, tmax, state, first, last):
        a = img.numpy()
        st = state.numpy()
        tmin, tmax = a.dtype.type(tmin), a.dtype.type(tmax)
        if first:
            mv = a[0].copy()
            start = (a[0] >= tmin) & (a[0] <= tmax)
            done = np.zeros(mv.shape, bool)
        else:
            mv = st[0].astype(a.dtype)
            start = (st[2] & 1).astype(bool)
            done = (st[2] & 2).astype(bool)
        for z in range(a.shape[0]):
            v = a[z]
            act = ~done
            gt = v > mv
            stop = act & ~gt & (v < mv) & start
            mv = np.where(act & gt, v, mv)
            done |= stop
            start |= act & ~stop & (v >= tmin) & (v <= tmax)
        st[0] = mv.astype(np.int32)
        st[1] = 0
        st[2] = start.astype(np.int32) | (done.astype(np.int32) << 1)
        return torch.from_numpy(np.ascontiguousarray(mv)) if last else None

    def fcm(self, img, n, axis, wl, ww, tmip):
        a = img.numpy()
        out = np.zeros([(a.shape[1], a.shape[2]), (a.shape[0], a.shape[2]), (a.shape[0], a.shape[1])][axis], a.dtype)
        self.o.fast_countour_mip(a, n, axis, wl, ww, tmip, out)
        return torch.from_numpy(out)

    # ---- flood fill
    def ff_begin(self, data, out, seeds, t0, t1, fill, strct):
        d, o = data.numpy(), out.numpy()
        passable = (d >= t0) & (d <= t1) & (o != fill)
        reached = np.zeros(d.shape, bool)
        for (x, y, z) in seeds:
            if t0 <= d[z, y, x] <= t1:
                passable[z, y, x] = True
                reached[z, y, x] = True
        return dict(passable=passable, reached=reached, out=o, fill=fill, strct=np.ascontiguousarray(strct, np.uint8),
                    changed=0)

    def ff_converge(self, st):
        seeds = [(int(x), int(y), int(z)) for z, y, x in np.argwhere(st["reached"])]
        if not seeds:
            return
        grown = np.zeros(st["reached"].shape, np.uint8)
        self.o._floodfill_threshold_core(st["passable"].astype(np.uint8), seeds, 1, 1, 1, st["strct"], grown)
        st["reached"] |= grown.astype(bool)

    def ff_get_planes(self, st, zs):
        return torch.from_numpy(np.stack([st["reached"][z] for z in zs]).astype(np.int32))

    def ff_merge_planes(self, st, zs, planes):
        st["changed"] = 0
        for z, pl in zip(zs, planes):
            new = pl.numpy().astype(bool) & st["passable"][z] & ~st["reached"][z]
            if new.any():
                st["reached"][z] |= new
                st["changed"] = 1

    def ff_changed(self, st):
        return torch.tensor([st["changed"]], dtype=torch.int32)

    def ff_finish(self, st):
        st["out"][st["reached"]] = st["fill"]

    # ---- contour volume / projections of own planes (rays along y or x ride on the z-ray walkers)
    def fcm_volume(self, img, n, axis):
        return torch.from_numpy(self.o.fcm_volume(img.numpy(), n, axis))

    def _as_z(self, img, axis):
        return img.permute(1, 0, 2).contiguous() if axis == 1 else img.permute(2, 0, 1).contiguous()

    def mida(self, img, axis, wl, ww, minmax):
        t = self._as_z(img, axis)
        return self.mida_z(t, wl, ww, minmax, self.ray_state(t), True, True)

    def lmip(self, img, axis, tmin, tmax):
        t = self._as_z(img, axis)
        return self.lmip_z(t, tmin, tmax, self.ray_state(t), True, True)

    # ---- fill holes
    def fh_hist(self, mask, labels, nlabels):
        sizes = np.bincount(labels.numpy().view(np.uint32).ravel(), minlength=nlabels + 1).astype(np.uint32)
        return dict(mask=mask, labels=labels, sizes=torch.from_numpy(sizes.view(np.int32).copy()))

    def fh_sizes(self, st):
        return st["sizes"]

    def fh_apply(self, st, max_size):
        sizes = st["sizes"].numpy().view(np.uint32)
        ok = (sizes > 0) & (sizes <= max_size)
        if not ok.any():
            return False
        m = st["mask"].numpy()
        m[ok[st["labels"].numpy().view(np.uint32)]] = 254
        return True

    # ---- watershed: the NumPy model of the flood (tests/ws_model.py), staged like the C ABI
    def ws_preprocess(self, image_i16, use_ww_wl, wl, ww, global_min=None):
        from oracle import watershed as W
        a = image_i16.numpy()
        if use_ww_wl:
            return torch.from_numpy(W.get_LUT_value(a, ww, wl).astype("uint16").view(np.int16))
        return torch.from_numpy((a - np.int16(global_min)).astype("uint16").view(np.int16))

    def ws_local_min(self, image_i16):
        return int(image_i16.numpy().min())

    def ws_gradient(self, pre, size):
        from scipy import ndimage
        return torch.from_numpy(ndimage.morphological_gradient(pre.numpy().view(np.uint16), size).view(np.int16))

    def ws_begin(self, cost_u16, markers_i16, mode, frozen_lo, frozen_hi):
        import ws_model
        return dict(m=ws_model.Model(cost_u16.numpy().view(np.uint16), markers_i16.numpy(), mode, frozen_lo, frozen_hi))

    def ws_converge(self, st, what):
        st["m"].converge_cost() if what == 0 else st["m"].converge_labels()

    def ws_label_begin(self, st):
        st["m"].label_begin()

    def ws_get_plane(self, st, what, z):
        if what == 0:
            return torch.from_numpy(st["m"].get_plane(0, z).astype(np.uint32).view(np.uint8).reshape(-1).copy())
        k, s = st["m"].get_plane(1, z)
        return torch.from_numpy(np.concatenate([k.astype(np.uint64).view(np.uint8).reshape(-1),
                                                s.astype(np.uint16).view(np.uint8).reshape(-1)]))

    def ws_merge_plane(self, st, what, z, plane):
        m = st["m"]
        ny, nx = m.shape[1:]
        b = plane.numpy()
        if what == 0:
            return int(m.merge_plane(0, z, b.view(np.uint32).reshape(ny, nx)))
        k = b[: ny * nx * 8].view(np.uint64).reshape(ny, nx)
        s = b[ny * nx * 8:].view(np.uint16).reshape(ny, nx).astype(np.int64)
        return int(m.merge_plane(1, z, (k, s)))

    def ws_finish(self, st, want_ambiguous):
        lab, amb = st["m"].labels()
        return torch.from_numpy(lab), (torch.from_numpy(amb) if want_ambiguous else None)

    # ---- marching cubes
    def mc_count(self, vol, iso, skip_last):
        a = vol.numpy()
        inside = a >= iso
        cx = np.zeros(a.shape, bool); cy = cx.copy(); cz = cx.copy()
        cx[:, :, :-1] = inside[:, :, :-1] != inside[:, :, 1:]
        cy[:, :-1, :] = inside[:, :-1, :] != inside[:, 1:, :]
        cz[:-1] = inside[:-1] != inside[1:]
        nv = cx.astype(np.int64) + cy + cz
        voff_full = np.cumsum(nv.ravel()) - nv.ravel()
        V_full = int(nv.sum())
        V_own = int(voff_full.reshape(a.shape)[-1, 0, 0]) if skip_last else V_full
        _, T = self.o.marching_cubes(a, iso)
        return dict(vol=a, iso=iso, skip_last=skip_last, cx=cx, cy=cy, cz=cz,
                    voff=voff_full.reshape(a.shape), V=V_own, T=len(T), V_full=V_full)

    def mc_plane0_records(self, st):
        return torch.from_numpy(np.stack([st["voff"][0], st["cx"][0].astype(np.int64), st["cy"][0].astype(np.int64),
                                          st["cz"][0].astype(np.int64)]))

    def mc_emit(self, st, spacing, origin_index, flip_y, vbase, foreign, foreign_base):
        V, T = self.o.marching_cubes(st["vol"], st["iso"], spacing, origin_index, flip_y)
        T = T.copy()
        own = T < st["V"]
        out = np.where(own, T + vbase, 0)
        if (~own).any():
            f = foreign.numpy()
            nz, ny, nx = st["vol"].shape
            base_last = int(st["voff"][-1, 0, 0])
            # reverse lookup: which (y, x, axis) of the last plane is local vertex id i?
            lut = {}
            for y in range(ny):
                for x in range(nx):
                    k = int(st["voff"][-1, y, x])
                    for ax, c in enumerate((st["cx"][-1, y, x], st["cy"][-1, y, x])):  # no +z edge in the last plane
                        if c:
                            lut[k] = (y, x, ax)
                            k += 1
            for idx in np.argwhere(~own):
                i = int(T[tuple(idx)])
                assert i >= base_last
                y, x, ax = lut[i]
                rank_in_voxel = int(f[1, y, x]) if ax == 1 else 0
                out[tuple(idx)] = foreign_base + int(f[0, y, x]) + rank_in_voxel
        return torch.from_numpy(V[: st["V"]].copy()), torch.from_numpy(out.astype(np.int32))


def _setup(rank, world):
    from invesalius3_b200 import dist as d
    g = global_volume()
    shard = d.ZShard(g.shape[0], rank, world)
    return d, g, shard


def rank_halo_and_mip(rank, world, device):
    d, g, shard = _setup(rank, world)
    be = CpuBackend()
    ext = torch.from_numpy(ext_slab(g, shard).copy())
    want = ext.clone()
    if shard.has_lo: ext[0] = -7
    if shard.has_hi: ext[-1] = -7
    d.exchange_halo(ext, shard)
    assert torch.equal(ext, want)
    own = torch.from_numpy(np.ascontiguousarray(g[shard.z0:shard.z1]))
    res = {}
    for axis in (0, 1, 2):
        for kind in ("max", "min", "mean"):
            res[(axis, kind)] = d.mip(own, axis, kind, shard, backend=be).numpy()
    m = d.threshold(own, *THR, shard, backend=be).numpy()
    return res, (shard.z0, shard.z1, m)


def test_halo_mip_threshold_two_ranks(orc):
    out = run_ranks("rank_halo_and_mip", "test_dist_gloo")
    g = global_volume()
    for rank in (0, 1):
        res, (z0, z1, m) = out[rank]
        for (axis, kind), got in res.items():
            want = {"max": g.max, "min": g.min, "mean": g.mean}[kind](axis)
            assert got.dtype == want.dtype and np.array_equal(got, want), (rank, axis, kind)
        want = np.zeros(g.shape, np.uint8)
        orc.threshold(g, *THR, want, False)
        assert np.array_equal(m, want[z0:z1])


def rank_contour_mip(rank, world, device):
    d, g, shard = _setup(rank, world)
    be = CpuBackend()
    ext = torch.from_numpy(ext_slab(g, shard).copy())
    res = {}
    for axis in (0, 1, 2):
        for tmip in (0, 1, 2):
            res[(axis, tmip)] = d.fast_countour_mip(ext, 2.0, axis, 300, 600, tmip, shard, backend=be).numpy()
    return res


def test_contour_mip_two_ranks(orc):
    """Sharded contour-MIP (halo planes feed the central differences) = the whole-volume one."""
    out = run_ranks("rank_contour_mip", "test_dist_gloo")
    g = global_volume()
    for (axis, tmip), got in out[0].items():
        want = np.zeros([(g.shape[1], g.shape[2]), (g.shape[0], g.shape[2]), (g.shape[0], g.shape[1])][axis], np.int16)
        orc.fast_countour_mip(g, 2.0, axis, 300, 600, tmip, want)
        assert np.array_equal(got, want), (axis, tmip)
        assert np.array_equal(out[1][(axis, tmip)], want), (axis, tmip)


def rank_rays_along_z(rank, world, device):
    d, g, shard = _setup(rank, world)
    be = CpuBackend()
    own = torch.from_numpy(np.ascontiguousarray(g[shard.z0:shard.z1]))
    res = {}
    for wl, ww in ((300, 600), (40, 1), (3000, 30000)):
        res[("mida", wl, ww)] = d.mida(own, 0, wl, ww, shard, backend=be).numpy()
    for tmin, tmax in ((700, 3033), (-200, 100)):
        res[("lmip", tmin, tmax)] = d.lmip(own, 0, tmin, tmax, shard, backend=be).numpy()
    return res


def test_mida_lmip_along_z_two_ranks(orc):
    """Rays that cross the shards: the state hand-off reproduces the whole-volume walk."""
    out = run_ranks("rank_rays_along_z", "test_dist_gloo")
    g = global_volume()
    for key, got in out[0].items():
        want = np.zeros(g.shape[1:], np.int16)
        if key[0] == "mida":
            orc.mida(g, 0, key[1], key[2], want)
        else:
            orc.lmip(g, 0, key[1], key[2], want)
        assert np.array_equal(got, want), key
        assert np.array_equal(out[1][key], want), key


def test_mida_lmip_along_z_three_ranks(orc):
    """World size 3: the middle shard both receives and forwards the ray state."""
    out = run_ranks("rank_rays_along_z", "test_dist_gloo", world=3)
    g = global_volume()
    for key, got in out[0].items():
        want = np.zeros(g.shape[1:], np.int16)
        (orc.mida if key[0] == "mida" else orc.lmip)(g, 0, key[1], key[2], want)
        for r in range(3):
            assert np.array_equal(out[r][key], want), (key, r)


def test_contour_mip_three_ranks(orc):
    out = run_ranks("rank_contour_mip", "test_dist_gloo", world=3)
    g = global_volume()
    for (axis, tmip), got in out[1].items():   # the middle shard has a halo plane on both sides
        want = np.zeros([(g.shape[1], g.shape[2]), (g.shape[0], g.shape[2]), (g.shape[0], g.shape[1])][axis], np.int16)
        orc.fast_countour_mip(g, 2.0, axis, 300, 600, tmip, want)
        assert np.array_equal(got, want), (axis, tmip)
    from invesalius3_b200 import dist as d
    shard = d.ZShard(4, 0, 1)
    with pytest.raises(ValueError):    # LMIP bounds 700 / 3033 do not fit uint8: the reference panics
        d.fast_countour_mip(torch.zeros((4, 4, 4), dtype=torch.uint8), 2.0, 1, 30, 60, 1, shard, backend=CpuBackend())


def ff_cases(g):
    from scipy.ndimage import generate_binary_structure
    def first(z):
        yy, xx = np.nonzero((g[z] >= 100) & (g[z] <= 3071) & (np.arange(g.shape[1])[:, None] != 10))
        return (int(xx[0]), int(yy[0]), z)
    return ((generate_binary_structure(3, 1), [first(2)]),
            (generate_binary_structure(3, 3), [first(21), first(11), (0, 0, 0)]))


def rank_floodfill(rank, world, device):
    d, g, shard = _setup(rank, world)
    be = CpuBackend()
    results = []
    for strct, seeds in ff_cases(g):
        data = torch.from_numpy(ext_slab(g, shard))
        out_g = np.zeros(g.shape, np.uint8)
        out_g[:, 10, :] = 254                       # a pre-filled wall crossing the shard boundary
        out = torch.from_numpy(ext_slab(out_g, shard).copy())
        outer = d.floodfill_threshold(data, seeds, 100, 3071, 254, strct, out, shard, backend=be)
        results.append((shard.z0, shard.z1, shard.interior(out).numpy().copy(), outer))
    return results


def test_floodfill_two_ranks(orc):
    out = run_ranks("rank_floodfill", "test_dist_gloo")
    g = global_volume()
    for case, (strct, seeds) in enumerate(ff_cases(g)):
        want = np.zeros(g.shape, np.uint8); want[:, 10, :] = 254
        orc.floodfill_threshold(g, seeds, 100, 3071, 254, strct, want)
        got = np.concatenate([out[r][case][2] for r in (0, 1)])
        assert np.array_equal(got, want), case
        grown = (want == 254); grown[:, 10, :] = False
        z_split = out[0][case][1]
        assert grown[:z_split].sum() > 50 and grown[z_split:].sum() > 50   # the region crosses the boundary
        assert out[0][case][3] == out[1][case][3] >= (2 if case == 0 else 1)


def rank_mc(rank, world, device):
    d, g, shard = _setup(rank, world)
    be = CpuBackend()
    mask = ((g >= THR[0]) & (g <= THR[1])).astype(np.uint8) * 255
    vol = torch.from_numpy(ext_slab(mask, shard, lo=False, hi=True))
    v, t, vbase, tv, tt = d.marching_cubes(vol, 127, (0.5, 0.75, 1.5), (-1, -1, 3), True, shard, backend=be)
    return v.numpy(), t.numpy(), vbase, tv, tt


def test_marching_cubes_two_ranks(orc):
    out = run_ranks("rank_mc", "test_dist_gloo")
    g = global_volume()
    mask = ((g >= THR[0]) & (g <= THR[1])).astype(np.uint8) * 255
    V, T = orc.marching_cubes(mask, 127, (0.5, 0.75, 1.5), (-1, -1, 3), True)
    gv = np.concatenate([out[0][0], out[1][0]])
    gt = np.concatenate([out[0][1], out[1][1]])
    assert out[0][3] == len(V) and out[0][4] == len(T) and out[1][2] == len(out[0][0])
    assert np.array_equal(gv, V)
    assert np.array_equal(gt.astype(np.int64), T)


# ---- watershed over Z shards: the plane-exchange protocol reaches the single-volume fixed point
def ws_case():
    g = global_volume((23, 20, 45), seed=5)
    mk = np.zeros(g.shape, np.uint8)
    mk[3, 4:7, 5:9] = 1; mk[19, 12:15, 30:34] = 2; mk[11, 2:4, 40:43] = 2; mk[12, 15:18, 3:6] = 1
    return g, mk


def rank_watershed(rank, world, device):
    from scipy.ndimage import generate_binary_structure
    d, _, _ = _setup(rank, world)
    g, mk = ws_case()
    shard = d.ZShard(g.shape[0], rank, world)
    be = CpuBackend()
    st6 = generate_binary_structure(3, 1)
    res = {}
    for alg in ("Watershed", "Watershed IFT"):
        for ww_wl in (True, False):
            lab, amb, ex = d.watershed(torch.from_numpy(ext_slab(g, shard)), torch.from_numpy(ext_slab(mk, shard)), st6, alg,
                                       3, ww_wl, 300, 900, shard, backend=be, return_ambiguous=True)
            res[(alg, ww_wl)] = (lab.numpy().copy(), amb.numpy().copy(), ex)
    with pytest.raises(NotImplementedError):
        d.watershed(torch.from_numpy(ext_slab(g, shard)), torch.from_numpy(ext_slab(mk, shard)),
                    generate_binary_structure(3, 3), "Watershed", 3, True, 300, 900, shard, backend=be)
    return res


def ws_whole(g, mk, alg, ww_wl):
    """The same model on the whole volume (no shards), pre-processing as the reference's."""
    import ws_model
    from oracle import watershed as W
    from scipy import ndimage
    pre = W.preprocess(g, ww_wl, 300, 900)
    m = mk.astype("int16")
    if alg == "Watershed":
        pre = ndimage.morphological_gradient(pre, 3)
    elif not ww_wl:
        m = mk.astype("int8").astype("int16")
    return ws_model.flood(pre, m, 1 if alg == "Watershed" else 0)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_watershed_ranks(orc, world):
    out = run_ranks("rank_watershed", "test_dist_gloo", world=world)
    g, mk = ws_case()
    for alg in ("Watershed", "Watershed IFT"):
        for ww_wl in (True, False):
            lab = np.concatenate([out[r][(alg, ww_wl)][0] for r in range(world)])
            amb = np.concatenate([out[r][(alg, ww_wl)][1] for r in range(world)])
            want_lab, want_amb = ws_whole(g, mk, alg, ww_wl)
            assert np.array_equal(lab, want_lab), (alg, ww_wl)
            assert np.array_equal(amb, want_amb), (alg, ww_wl)
            assert len({out[r][(alg, ww_wl)][2] for r in range(world)}) == 1      # every rank saw the same exchanges
            assert out[0][(alg, ww_wl)][2] >= 3                                    # labels did cross the boundary


# ---- fill holes over Z shards: one all_reduce of the label sizes
def fh_case():
    from scipy import ndimage
    rng = np.random.default_rng(11)
    mask = (ndimage.gaussian_filter(rng.normal(size=(23, 20, 45)), 1.0) > 0.02).astype(np.uint8) * 255
    labels, nlabels = ndimage.label(mask == 0, ndimage.generate_binary_structure(3, 1), output=np.uint32)
    return mask, labels, int(nlabels)


def rank_fill_holes(rank, world, device):
    d, _, _ = _setup(rank, world)
    mask, labels, nlabels = fh_case()
    shard = d.ZShard(mask.shape[0], rank, world)
    res = {}
    for max_size in (0, 5, 40, 10 ** 6):
        m = torch.from_numpy(mask[shard.z0:shard.z1].copy())
        lab = torch.from_numpy(labels[shard.z0:shard.z1].view(np.int32).copy())
        ret = d.fill_holes_automatically(m, lab, nlabels, max_size, shard, backend=CpuBackend())
        res[max_size] = (ret, m.numpy().copy())
    return res


@pytest.mark.parametrize("world", [2, 3, 4])
def test_fill_holes_ranks(orc, world):
    out = run_ranks("rank_fill_holes", "test_dist_gloo", world=world)
    mask, labels, nlabels = fh_case()
    for max_size in (0, 5, 40, 10 ** 6):
        want = mask.copy()
        ret = orc.fill_holes_automatically(want, labels, nlabels, max_size)
        got = np.concatenate([out[r][max_size][1] for r in range(world)])
        assert np.array_equal(got, want), max_size
        assert all(out[r][max_size][0] == ret for r in range(world)), max_size


# ---- four ranks, uneven slabs (23 planes -> 6, 6, 6, 5): the same protocols across three boundaries
def test_floodfill_four_ranks(orc):
    world = 4
    out = run_ranks("rank_floodfill", "test_dist_gloo", world=world)
    g = global_volume()
    for case, (strct, seeds) in enumerate(ff_cases(g)):
        want = np.zeros(g.shape, np.uint8); want[:, 10, :] = 254
        orc.floodfill_threshold(g, seeds, 100, 3071, 254, strct, want)
        assert [out[r][case][0] for r in range(1, world)] == [out[r][case][1] for r in range(world - 1)]
        got = np.concatenate([out[r][case][2] for r in range(world)])
        assert np.array_equal(got, want), case
        grown = (want == 254); grown[:, 10, :] = False
        assert all(grown[out[r][case][0]:out[r][case][1]].any() for r in range(world))   # every shard takes part
        assert len({out[r][case][3] for r in range(world)}) == 1                          # one vote, one round count


def test_marching_cubes_four_ranks(orc):
    world = 4
    out = run_ranks("rank_mc", "test_dist_gloo", world=world)
    g = global_volume()
    mask = ((g >= THR[0]) & (g <= THR[1])).astype(np.uint8) * 255
    V, T = orc.marching_cubes(mask, 127, (0.5, 0.75, 1.5), (-1, -1, 3), True)
    gv = np.concatenate([out[r][0] for r in range(world)])
    gt = np.concatenate([out[r][1] for r in range(world)])
    assert all(out[r][3] == len(V) and out[r][4] == len(T) for r in range(world))
    assert [out[r][2] for r in range(world)] == np.cumsum([0] + [len(out[r][0]) for r in range(world - 1)]).tolist()
    assert np.array_equal(gv, V)
    assert np.array_equal(gt.astype(np.int64), T)
