"""
The gray-volume path: 3-D SLIC (csrc/slic3d.cu) against oracle/slic3d_oracle.c -- bit-exact label
volumes --, the 6-connected supervoxel graph, and pipe_gray3d_slic_features_model_graphcut (imsegm/pipelines.py:382-431).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from pyimsegm_b200.engine import get_engine
    return get_engine()


def _blobs(shape, seed, noise=0.08):
    rng = np.random.RandomState(seed)
    zz, yy, xx = np.mgrid[:shape[0], :shape[1], :shape[2]]
    vol = 0.3 + 0.4 * ((xx > shape[2] // 2) ^ (yy > shape[1] // 3)) + 0.15 * (zz > shape[0] // 2)
    return np.clip(vol + rng.normal(0, noise, shape), 0, 1)


@pytest.mark.parametrize('case', ['doctest', 'aniso', 'iso', 'thin', 'u8', 'f32'])
def test_volume_slic_label_map_bit_exact(oracle, case):
    from pyimsegm_b200 import superpixels as sp
    if case == 'doctest':       # imsegm/superpixels.py:81-86
        np.random.seed(0)
        vol, args = np.random.random((100, 100, 10)), (20, 0.2, (1, 1, 5))
    elif case == 'aniso':       # the shape and spacing of pipe_gray3d's doctest (pipelines.py:404-408)
        vol, args = _blobs((5, 125, 150), 1), (15, 0.2, (12, 1, 1))
    elif case == 'iso':
        vol, args = _blobs((24, 40, 36), 2), (8, 0.3, (1, 1, 1))
    elif case == 'thin':
        vol, args = _blobs((2, 61, 47), 3, noise=0.2), (9, 0.15, (3, 1, 1))
    elif case == 'u8':
        vol, args = (_blobs((12, 50, 44), 4) * 255).astype(np.uint8), (10, 0.3, (2, 1, 1))
    else:
        vol, args = _blobs((9, 33, 70), 5).astype(np.float32), (7, 0.25, (1, 1, 2))
    got = sp.segment_slic_img3d_gray(vol, *args)
    want = oracle.segment_slic_img3d_gray(vol, *args)
    assert got.shape == vol.shape and got.dtype == np.int64
    assert np.array_equal(got, want)
    assert set(np.unique(got)) == set(range(got.max() + 1))


def test_volume_kmeans_and_connectivity_separately(oracle, eng):
    """the sweeps without the connectivity pass, and the connectivity pass alone on label volumes that force the oversize split
    (small max_size) and long merge chains (large min_size)"""
    import ctypes as C
    from pyimsegm_b200 import _lib
    vol = _blobs((10, 48, 52), 7, noise=0.15)
    n_seg, compact, spacing = 60, 3, (2, 1, 1)
    km, _ = eng.slic3d(eng.to_device(vol, 'volume'), n_seg, compact, spacing, enforce_connectivity=False)
    km = eng.to_host(km).copy()
    want_km = oracle.slic3d(vol, n_seg, compact, spacing, return_kmeans=True)
    assert np.array_equal(km, want_km)
    rng = np.random.RandomState(8)
    noisy = want_km.copy()
    flip = rng.rand(*noisy.shape) < 0.15
    noisy[flip] = rng.randint(0, noisy.max() + 1, flip.sum())           # speckle: many tiny components
    D, H, W = noisy.shape
    for min_size, max_size in ((3, 40), (30, 200), (200, 100000), (1, 5)):
        want = np.empty_like(noisy)
        n = oracle.lib().oracle_enforce_connectivity3d(noisy.ctypes.data_as(C.POINTER(C.c_int64)), D, H, W, C.c_long(min_size), C.c_long(max_size),
                                                       want.ctypes.data_as(C.POINTER(C.c_int64)))
        d_in = eng.to_device(noisy.astype(np.int32), 'conn3d_in')
        out = eng.buf('conn3d_out', (D, H, W), eng.torch.int32)
        nl = eng.buf('conn3d_n', (1,), eng.torch.int32)
        wsb = eng.lib.isb_connectivity3d_workspace_bytes(D, H, W, max_size)
        ws = eng.buf('conn3d_ws', (wsb,), eng.torch.uint8)
        _lib.check(eng.lib.isb_enforce_connectivity3d(_lib.ptr(d_in), D, H, W, min_size, max_size, _lib.ptr(out), _lib.ptr(nl), _lib.ptr(ws),
                                                      C.c_size_t(wsb), _lib.stream_ptr()))
        assert np.array_equal(eng.to_host(out), want), (min_size, max_size)
        assert int(eng.to_host(nl)[0]) == max(n, 1)


def test_volume_graph_and_pipeline(oracle):
    from pyimsegm_b200 import graph_cuts as gc
    from pyimsegm_b200 import pipelines as pl
    from pyimsegm_b200 import superpixels as sp
    # the doctest graph of make_graph_segm_connect_grid3d_conn6 (superpixels.py:186-194) through the device path
    grid_2d = np.array([[0] * 5 + [1] * 5, [2] * 5 + [3] * 5])
    grid = np.array([grid_2d, grid_2d + 4])
    edges, weights = gc.compute_edge_weights(grid, proba=np.eye(8), edge_type='model')
    assert sorted(map(tuple, edges.tolist())) == sorted(map(tuple, sp.make_graph_segm_connect_grid3d_conn6(grid)[1]))
    assert weights.shape == (12, ) and np.all((weights >= 1e-3) & (weights <= 1e3))
    # pipelines.py:404-408
    np.random.seed(0)
    image = np.random.random((5, 125, 150)) / 2.
    image[:, :, :75] += 0.5
    segm = pl.pipe_gray3d_slic_features_model_graphcut(image, 2, {'color': ['mean']})
    assert segm.shape == (5, 125, 150)
    left, right = segm[:, :, :70], segm[:, :, 80:]
    assert {np.bincount(left.ravel()).argmax(), np.bincount(right.ravel()).argmax()} == {0, 1}
    assert (left == np.bincount(left.ravel()).argmax()).mean() > 0.9 and (right == np.bincount(right.ravel()).argmax()).mean() > 0.9
    # every stage against the oracle given the same supervoxels: features and graph
    slic = sp.segment_slic_img3d_gray(image, sp_size=15, relative_compact=0.2, space=(12, 1, 1))
    assert np.array_equal(slic, oracle.segment_slic_img3d_gray(image, 15, 0.2, (12, 1, 1)))
    e_dev, _ = gc.compute_edge_weights(slic, proba=np.ones((slic.max() + 1, 2)) / 2, edge_type='spatial')
    assert sorted(map(tuple, e_dev.tolist())) == sorted(map(tuple, np.asarray(sp.make_graph_segm_connect_grid3d_conn6(slic)[1]).tolist()))
