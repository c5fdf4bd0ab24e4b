"""CPU tests: the C-ABI library loads and exports every symbol include/imsegm_b200.h declares (no compute calls
without a GPU), the host-side logic of the reference-API mirror, and the loud failure without a device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built_lib():
    from pyimsegm_b200 import build
    build.build()
    from pyimsegm_b200 import _lib
    return _lib


def test_header_symbols_all_exported(built_lib):
    header = open(os.path.join(ROOT, 'include', 'imsegm_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(isb_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 20
    handle = built_lib.lib()
    for name in declared:
        assert hasattr(handle, name), 'library does not export %s' % name
    assert declared == set(built_lib.SIGNATURES), 'ctypes binding table and header disagree'
    assert handle.isb_abi_version() >= 2
    assert handle.isb_profile_stage_count() == 12
    assert handle.isb_launch_count() == 0


def test_argument_validation_needs_no_gpu(built_lib):
    lib = built_lib.lib()
    rc = lib.isb_slic_prepare(None, 3, 4, 4, 3, None, 0, 1.0, 1, None, None, None)
    assert rc == built_lib.ISB_ERR_ARG and b'null' in lib.isb_last_error()
    with pytest.raises(ValueError):
        built_lib.check(rc)
    assert lib.isb_slic_kmeans_workspace_bytes(2048, 2048, 4900, 29, 29) > 0
    assert lib.isb_connectivity_workspace_bytes(64, 64) > 64 * 64 * 4
    assert lib.isb_alpha_expansion_workspace_bytes(5000, 3, 15000) > 0


def test_no_cpu_fallback_without_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a device is present')
    from pyimsegm_b200 import superpixels
    with pytest.raises(built_lib.NativeLibraryError):
        superpixels.segment_slic_img2d(np.zeros((32, 32, 3)), 8, 0.2)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'pyimsegm_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(import|from)\s+oracle\b', src, flags=re.M), f
                assert 'liboracle' not in src, f


def test_seed_grid_and_parameter_mapping(oracle):
    from pyimsegm_b200.engine import gaussian_half_kernel, regular_grid_steps, slic_seed_grid
    from pyimsegm_b200.superpixels import slic_params
    assert slic_params((2048, 2048), 29, 0.2) == (4987, (29 * 0.2) ** 1.5)     # superpixels.py:57-58
    for shape, n in (((1, 100, 150), 37), ((1, 512, 512), 419), ((1, 2048, 2048), 4987), ((1, 7, 300), 10)):
        assert regular_grid_steps(shape, n) == oracle.regular_grid(shape, n)
    s, ty, tx = slic_seed_grid(512, 512, 419)
    s_o, ty_o, tx_o = oracle.slic_seeds(512, 512, 419)
    assert np.array_equal(s, s_o) and (ty, tx) == (ty_o, tx_o)
    w, r = gaussian_half_kernel(1.0)
    w_o, r_o = oracle.gaussian_weights(1.0)
    assert r == r_o == 4 and np.array_equal(w, w_o)


def test_host_side_graph_helpers_match_reference_doctests():
    from pyimsegm_b200 import graph_cuts as gc
    from pyimsegm_b200 import superpixels as sp
    np.testing.assert_allclose(gc.create_pairwise_matrix_uniform(0.2, 3), [[0, .2, .2], [.2, 0, .2], [.2, .2, 0]])
    np.testing.assert_allclose(gc.create_pairwise_matrix_specif([((1, 2), 0.5), ((1, 0), 0.7)], 4),
                               [[0., 0.7, 1., 1.], [0.7, 0., 0.5, 1.], [1., 0.5, 0., 1.], [1., 1., 1., 0.]])
    np.testing.assert_allclose(gc.create_pairwise_matrix([((1, 2), 0.5), ((0, 2), 0.7)], 3), [[0., 1., 0.7], [1., 0., 0.5], [0.7, 0.5, 0.]])
    trans = np.array([[25., 5., 0.], [5., 10., 8.], [0., 8., 30.]])
    np.testing.assert_allclose(np.round(gc.compute_pairwise_cost_from_transitions(trans), 3),
                               [[0.182, 1.526, 20.723], [1.526, 0.833, 1.056], [20.723, 1.056, 0.236]])
    np.testing.assert_allclose(np.round(gc.compute_pairwise_cost_from_transitions(np.eye(3)), 2),
                               [[0., 20.72, 20.72], [20.72, 0., 20.72], [20.72, 20.72, 0.]])
    assert sp.get_neighboring_segments([[0, 1], [1, 2], [1, 3], [2, 3]]) == [[1], [0, 2, 3], [1, 3], [1, 2]]
    v, e = sp.make_graph_segment_connect_edges(np.arange(4), np.array([[0, 1], [1, 0], [2, 2], [3, 1]]))
    assert [list(map(int, x)) for x in e] == [[0, 1], [1, 3]]
    with pytest.raises(ValueError):
        gc.create_pairwise_matrix(np.ones((2, 2)), 3)
    un = gc.compute_unary_cost(np.array([[0.0, 1.0], [0.5, 0.5]]))
    np.testing.assert_allclose(un, np.abs(np.log([[0.01, 0.99], [0.5, 0.5]])))
    from pyimsegm_b200.utilities import ImageDimensionError
    from pyimsegm_b200 import descriptors as ds
    with pytest.raises(ImageDimensionError):
        ds._check_color_image(np.zeros((20, 25, 1)))
    with pytest.raises(ImageDimensionError):
        ds._check_color_image_segm(np.zeros((12, 15, 3)), np.zeros((15, 12)))
    assert ds._check_unrecognised_feature_group({'color': [], 'texture': []}) == ['texture']
    assert ds._check_unrecognised_feature_names(['mean', 'average']) == ['average']


def test_drop_in_alias_package():
    import imsegm.pipelines as pl
    import imsegm.graph_cuts as gc
    assert pl.pipe_color2d_slic_features_model_graphcut.__module__ == 'pyimsegm_b200.pipelines'
    assert gc.MIN_MAX_EDGE_WEIGHT == 1e3 and gc.MIN_UNARY_PROB == 0.01 and gc.MAX_PAIRWISE_COST == 1e5
    with pytest.raises(ValueError):
        pl.compute_color2d_superpixels_features(np.zeros((32, 32, 3)), {'color': ['mean']}, sp_regul=0.)


def test_color_space_shim_matches_reference_doctest():
    """imsegm/descriptors.py:1227-1231: statistics of the HSV image (host conversion + host numpy statistics)"""
    from pyimsegm_b200 import descriptors as ds
    from pyimsegm_b200.color import convert_img_color_from_rgb
    image = np.zeros((2, 10, 3))
    image[:, 2:6, 0] = 1
    image[:, 3:7, 1] = 3
    image[:, 4:9, 2] = 2
    segm = np.array([[0] * 5 + [1] * 5] * 2)
    hsv = convert_img_color_from_rgb(image, 'hsv')
    got = np.round(np.hstack([ds.numpy_img2d_color_mean(hsv, segm), ds.numpy_img2d_color_std(hsv, segm)]), 3)
    assert got.tolist() == [[0.139, 0.533, 1.4, 0.176, 0.452, 1.356], [0.439, 0.733, 2., 0.244, 0.389, 1.095]]
    assert convert_img_color_from_rgb(np.ones((50, 75, 3)), 'hsv').shape == (50, 75, 3)
    assert convert_img_color_from_rgb(np.ones((5, 7, 3)), 'unknown').shape == (5, 7, 3)
    rng = np.random.RandomState(0)
    rgb = rng.random_sample((6, 9, 3))
    lab = convert_img_color_from_rgb(rgb, 'lab')
    assert 0 <= lab[..., 0].min() and lab[..., 0].max() <= 100
    np.testing.assert_allclose(convert_img_color_from_rgb(np.full((2, 2, 3), 1.0), 'xyz')[0, 0], [0.950456, 1.0, 1.088754], rtol=1e-6)
