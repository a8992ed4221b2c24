"""CPU checker of context-aware smoothing (oracle/mesh.c: mesh.rs:27-395 restated sequentially, the
reference's quirks kept; the reference has no test for it — parity unpinned): behaviour the restatement
must show."""
import numpy as np
import pytest


def _sphere_mesh(orc, n=24, r=8):
    zz, yy, xx = np.indices((n, n, n))
    vol = (((zz - n / 2) ** 2 + (yy - n / 2) ** 2 + (xx - n / 2) ** 2) < r * r).astype(np.uint8) * 255
    V, F = orc.marching_cubes(vol, 127, (1.0, 1.0, 1.0), (0, 0, 0), False)
    F4 = np.ascontiguousarray(np.c_[np.full(len(F), 3), F].astype(np.int64))
    a, b, c = V[F[:, 0]].astype(np.float64), V[F[:, 1]].astype(np.float64), V[F[:, 2]].astype(np.float64)
    nrm = np.cross(b - a, c - a)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    return V.copy(), F4, np.ascontiguousarray(nrm.astype(np.float32))


def test_every_referenced_vertex_is_a_seed_and_the_staircase_flattens(orc):
    """find_staircase_artifacts' max / min tracking flags a vertex at its first face (min is still
    f64::MAX there), so every vertex that has a face gets distance 0 and weight (1 - 0) * (1 - bmin) +
    bmin; the smoothing is then a uniform Taubin pass: the voxel staircase of a sphere gets rounder."""
    V, F4, N = _sphere_mesh(orc)
    v = V.copy()
    w = orc.ca_smoothing(v, F4, N, 0.7, 3.0, 0.1, 10, return_weights=True)
    assert np.array_equal(w, np.full(len(V), (1.0 - 0.0) * (1.0 - 0.1) + 0.1))
    rad0 = np.linalg.norm(V - 12.0, axis=1)
    rad1 = np.linalg.norm(v - 12.0, axis=1)
    assert rad1.std() < 0.6 * rad0.std() and abs(rad1.mean() - rad0.mean()) < 0.15     # smoother, no shrinkage (Taubin)
    assert v.dtype == np.float32 and np.isfinite(v).all()


def test_unreferenced_vertices_keep_bmin_and_their_place(orc):
    V, F4, N = _sphere_mesh(orc)
    V2 = np.concatenate([V, np.array([[100, 100, 100]], np.float32)])
    v = V2.copy()
    w = orc.ca_smoothing(v, F4, N, 0.7, 3.0, 0.25, 3, return_weights=True)
    assert w[-1] == 0.25 and np.array_equal(v[-1], V2[-1])
    v0 = V.copy(); orc.ca_smoothing(v0, F4, N, 0.7, 3.0, 0.25, 3)
    assert np.array_equal(v[:-1], v0)
    with pytest.raises(IndexError):
        bad = F4.copy(); bad[0, 1] = len(V2) + 1
        orc.ca_smoothing(V2.copy(), bad, N, 0.7, 3.0, 0.25, 1)


def test_zero_iterations_and_translation_equivariance(orc):
    V, F4, N = _sphere_mesh(orc)
    v = V.copy(); orc.ca_smoothing(v, F4, N, 0.7, 3.0, 0.1, 0)
    assert np.array_equal(v, V)
    a = V.copy(); orc.ca_smoothing(a, F4, N, 0.7, 3.0, 0.1, 4)
    shift = np.array([64.0, -32.0, 16.0], np.float32)           # exactly representable offsets
    b = (V + shift).astype(np.float32); orc.ca_smoothing(b, F4, N, 0.7, 3.0, 0.1, 4)
    assert np.abs((b - shift) - a).max() < 1e-4
