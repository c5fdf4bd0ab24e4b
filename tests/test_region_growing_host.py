"""Host-side pieces of pyimsegm_b200.region_growing against the reference's doctest values (imsegm/region_growing.py; the line of
each golden is cited).  Nothing here touches the GPU: these functions are numpy bookkeeping around the device calls."""
import numpy as np


def _rg():
    from pyimsegm_b200 import region_growing
    return region_growing


def test_cumulative_distrib_golden():
    """region_growing.py:344-348"""
    cdist = _rg().compute_cumulative_distrib(np.array([[1, 2]]), np.array([[1.5, 0.5], [0.5, 1]]), np.array([0.5]), 6)
    np.testing.assert_allclose(np.round(cdist, 2), [[1., 0.67, 0.34, 0.12, 0.03, 0., 0.], [1., 0.98, 0.5, 0.02, 0., 0., 0.]])


def test_cdf_histograms_golden():
    """region_growing.py:564-570"""
    list_rays = [[9, 4, 9], [4, 9, 7], [9, 7, 11], [10, 8, 10], [9, 11, 8], [4, 8, 5], [8, 10, 6], [9, 7, 11]]
    chist = _rg().transform_rays_model_cdf_histograms(list_rays, nb_bins=5)
    assert chist == [[1.0, 1.0, 1.0, 1.0, 0.75, 0.75, 0.75, 0.625, 0.625, 0.0, 0.0, 0.0],
                     [1.0, 1.0, 1.0, 1.0, 0.875, 0.875, 0.875, 0.375, 0.25, 0.25, 0.0, 0.0],
                     [1.0, 1.0, 1.0, 1.0, 1.0, 0.75, 0.625, 0.5, 0.375, 0.375, 0.0, 0.0]]


def test_shape_prior_table_golden():
    """region_growing.py:601-620 (the reference evaluates a scipy interp2d object per point; here one bilinear lookup)"""
    rg = _rg()
    chist = [[1.0, 1.0, 0.8, 0.7, 0.6, 0.5, 0.3, 0.0, 0.0], [1.0, 1.0, 0.9, 0.8, 0.7, 0.3, 0.2, 0.2, 0.0],
             [1.0, 1.0, 1.0, 0.7, 0.6, 0.5, 0.3, 0.1, 0.1], [1.0, 1.0, 0.6, 0.5, 0.4, 0.3, 0.2, 0.0, 0.0]]
    centre = (1, 1)
    f = rg.compute_shape_prior_table_cdf
    assert f([1, 1], chist, centre) == 1.0
    assert f([10, 10], chist, centre) == 0.0
    assert abs(f([10, -10], chist, centre) - 0.100) < 1e-3
    assert abs(f([2, 3], chist, centre) - 0.805) < 1e-3
    assert abs(f([-3, -2], chist, centre) - 0.381) < 1e-3
    assert abs(f([3, -2], chist, centre) - 0.676) < 1e-3
    assert abs(f([2, 3], chist, centre, angle_shift=270) - 0.891) < 1e-3
    pts = np.array([[1, 1], [10, 10], [2, 3], [-3, -2], [3, -2]])
    many = rg.shape_prior_table_cdf_points(pts, chist, centre)
    np.testing.assert_allclose(many, [f(p, chist, centre) for p in pts])


def test_centre_moment_golden():
    """region_growing.py:710-721"""
    f = _rg().compute_centre_moment_points
    c, t = f(list(zip([0] * 10, np.arange(10))) + [(0, 0)] * 5)
    np.testing.assert_allclose(c, [0., 3.]); assert t == 0.0
    c, t = f(list(zip(np.arange(10), [0] * 10)) + [(10, 0)])
    np.testing.assert_allclose(c, [5., 0.]); assert t == 90.0
    c, t = f(list(zip(-np.arange(10), -np.arange(10))) + [(0, 0)] * 5)
    np.testing.assert_allclose(c, [-3., -3.]); assert t == 45.0
    c, t = f(list(zip(-np.arange(10), np.arange(10))) + [(-10, 10)])
    np.testing.assert_allclose(c, [-5., 5.]); assert t == 135.0


def test_update_shape_costs_table_cdf_golden():
    """region_growing.py:784-807"""
    rg = _rg()
    cdf = np.zeros((8, 20))
    cdf[:10] = 0.5
    cdf[:4] = 1.0
    points = np.array([[13, 16], [1, 5], [10, 15], [15, 25], [10, 5]])
    labels = np.ones(len(points))
    s_costs = np.zeros((len(points), 2))
    s_costs, centres, shifts, _ = rg.compute_update_shape_costs_points_table_cdf(s_costs, points, labels, [(0, 0)], [(np.inf, np.inf)], [0], [0],
                                                                                (None, cdf))
    assert centres.tolist() == [[10, 13]]
    assert shifts.tolist() == [209.]
    np.testing.assert_allclose(np.round(s_costs, 3), [[0., 0.673], [0., -0.01], [0., 0.184], [0., 0.543], [0., 0.374]])
    thr = dict(rg.RG2SP_THRESHOLDS)
    thr['centre_init'] = 1
    _, centres, _, _ = rg.compute_update_shape_costs_points_table_cdf(s_costs, points, labels, [(7, 18)], [(np.inf, np.inf)], [0], [0], (None, cdf),
                                                                      dict_thresholds=thr)
    np.testing.assert_allclose(np.round(centres, 1), [[7.5, 17.1]])


def test_pairwise_penalty_and_candidates_golden():
    """region_growing.py:1074-1077, 1098-1101"""
    rg = _rg()
    edges = np.array([[0, 1], [1, 2], [0, 3], [2, 3], [2, 4]])
    labels = np.array([0, 0, 1, 2, 1])
    np.testing.assert_allclose(rg.compute_pairwise_penalty(edges, labels, 0.05, 0.01), [0., 2.99573227, 2.99573227, 4.60517019, 0.])
    assert rg.get_neighboring_candidates([[1], [0, 2, 3], [1, 3], [1, 2]], np.array([0, 0, 1, 1]), 1) == [1]


def test_module_is_aliased():
    import imsegm.region_growing as alias
    from pyimsegm_b200 import region_growing
    assert alias.region_growing_shape_slic_graphcut is region_growing.region_growing_shape_slic_graphcut


def test_legacy_rotation_ray_tracer_golden():
    """descriptors.py:1562-1571 compute_ray_features_segm_2d_vectors (scipy shift / rotate like the original; host only)"""
    from pyimsegm_b200 import descriptors as ds
    seg = np.ones((100, 100), dtype=bool)
    yy, xx = np.mgrid[:100, :100]
    seg[(yy - 45) ** 2 + (xx - 55) ** 2 < 30 ** 2] = False
    assert ds.compute_ray_features_segm_2d_vectors(seg, (50, 50), 45).tolist() == [35, 29, 25, 23, 24, 29, 34, 36]
    assert ds.compute_ray_features_segm_2d_vectors(seg, (60, 40), 30, smooth_coef=1).tolist() == [35, 27, 18, 12, 10, 9, 12, 18, 27, 37, 45, 49]
    assert ds.compute_ray_features_segm_2d_vectors(seg, (40, 60), 20).tolist() == [25, 27, 29, 32, 34, 35, 37, 36, 36, 34, 32, 29, 27, 25, 24, 23, 24, 24]
