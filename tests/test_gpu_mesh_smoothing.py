"""Context-aware smoothing on the device against the sequential CPU restatement (oracle/mesh.c)."""
import numpy as np
import pytest
from scipy import ndimage

pytestmark = pytest.mark.gpu


def _mesh(orc, shape=(40, 48, 56), seed=1, spacing=(0.9, 0.9, 1.5)):
    rng = np.random.default_rng(seed)
    vol = (ndimage.gaussian_filter(rng.normal(size=shape), 2.0) > 0).astype(np.uint8) * 255
    vol[0] = vol[-1] = 0; vol[:, 0] = vol[:, -1] = 0; vol[:, :, 0] = vol[:, :, -1] = 0
    V, F = orc.marching_cubes(vol, 127, spacing, (0, 0, 0), True)
    F4 = np.ascontiguousarray(np.c_[np.full(len(F), 3), F].astype(np.int64))
    a, b, c = V[F[:, 0]].astype(np.float64), V[F[:, 1]].astype(np.float64), V[F[:, 2]].astype(np.float64)
    n = np.cross(b - a, c - a)
    n /= np.maximum(np.linalg.norm(n, axis=1), 1e-30)[:, None]
    return V.copy(), F4, np.ascontiguousarray(n.astype(np.float32))


@pytest.mark.parametrize("params", [(0.7, 3.0, 0.1, 10), (0.35, 1.5, 0.3, 3), (2.5, 3.0, 0.0, 1)])
def test_equals_sequential_restatement(orc, params):
    from invesalius3_b200 import mesh_ops
    V, F4, N = _mesh(orc)
    want = V.copy()
    orc.ca_smoothing(want, F4, N, *params)
    got = V.copy()
    mesh_ops.context_aware_smoothing(got, F4, N, *params)
    assert np.array_equal(got, want)
    assert np.abs(got - V).max() > 1e-3            # it did move the surface
    # Mesh object form, int32 faces, float64 normals (the reference's other dtype arms)
    m = mesh_ops.Mesh(vertices=V.copy(), faces=F4.astype(np.int32), normals=N.astype(np.float64))
    mesh_ops.ca_smoothing(m, *params)
    assert np.array_equal(m.vertices, want)


def test_unreferenced_vertices_and_errors(orc):
    from invesalius3_b200 import mesh_ops
    V, F4, N = _mesh(orc, (20, 24, 28), seed=2)
    V2 = np.concatenate([V, np.array([[1e3, 1e3, 1e3], [2e3, 0, 0]], np.float32)])     # two vertices no face uses
    want = V2.copy(); orc.ca_smoothing(want, F4, N, 0.7, 3.0, 0.1, 5)
    got = V2.copy(); mesh_ops.context_aware_smoothing(got, F4, N, 0.7, 3.0, 0.1, 5)
    assert np.array_equal(got, want) and np.array_equal(got[-2:], V2[-2:])
    bad = F4.copy(); bad[5, 2] = len(V) + 7
    with pytest.raises(ValueError):
        mesh_ops.context_aware_smoothing(V.copy(), bad, N, 0.7, 3.0, 0.1, 1)
    with pytest.raises(TypeError):
        mesh_ops.context_aware_smoothing(V.astype(np.float64), F4, N, 0.7, 3.0, 0.1, 1)
