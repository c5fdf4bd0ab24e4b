"""GPU parity of pyimsegm_b200.region_growing (BASELINE config 4) against the reference's doctest values
(imsegm/region_growing.py, the line of each golden is cited): every max-flow below is the device alpha-expansion, the graphs,
centres, histograms and Ray features come from the device kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rg():
    from pyimsegm_b200 import region_growing
    return region_growing


def _block_slic(h=15, w=20, step=2):
    slic = np.zeros((h, w), dtype=int)
    for i in range(int(np.ceil(h / float(step)))):
        for j in range(int(np.ceil(w / float(step)))):
            slic[i * step:(i + 1) * step, j * step:(j + 1) * step] = int(i * (w / step) + j)
    return slic


CHIST4 = [[1.] * 3 + [0.8, 0.7, 0.6, 0.5, 0.3, 0.1, 0.0], [1.] * 3 + [0.9, 0.8, 0.7, 0.3, 0.2, 0.2, 0.1],
          [1.] * 3 + [1.0, 0.7, 0.6, 0.5, 0.3, 0.1, 0.1], [1.] * 3 + [0.6, 0.5, 0.4, 0.3, 0.2, 0.1, 0.0]]


def _box(h, w, rows, cols):
    m = np.zeros((h, w), dtype=int)
    m[rows[0]:rows[1], cols[0]:cols[1]] = 1
    return m


def test_object_segmentation_graphcut_slic_golden():
    """region_growing.py:70-76"""
    rg = _rg()
    slic = np.array([[0] * 3 + [1] * 3 + [2] * 3 + [3] * 3 + [4] * 3, [5] * 3 + [6] * 3 + [7] * 3 + [8] * 3 + [9] * 3])
    segm = np.array([[0] * 15, [1] * 12 + [0] * 3])
    want = [0, 0, 0, 0, 0, 1, 1, 1, 1, 0]
    got = rg.object_segmentation_graphcut_slic(slic, segm, [(1, 7)], gc_regul=0., edge_coef=1., coef_shape=1.)
    assert got.dtype == np.int32 and got.tolist() == want
    dbg = {}
    assert rg.object_segmentation_graphcut_slic(slic, segm, [(1, 7)], gc_regul=1., edge_coef=1., debug_visual=dbg).tolist() == want
    assert len(dbg['unary_imgs']) == 2 and dbg['unary_imgs'][0].shape == slic.shape


def test_object_segmentation_graphcut_pixels_golden():
    """region_growing.py:182-200 (gco.cut_grid_graph -> the device alpha-expansion over the pixel grid)"""
    rg = _rg()
    segm = np.array([[0] * 10, [1] * 5 + [0] * 5, [1] * 4 + [0] * 6, [0] * 6 + [1] * 4, [0] * 5 + [1] * 5, [0] * 10])
    centres = [(1, 2), (4, 8)]
    got = rg.object_segmentation_graphcut_pixels(segm, centres, gc_regul=0., coef_shape=0.5)
    assert got.tolist() == [[0] * 10, [2, 2, 1, 2, 2, 0, 0, 0, 0, 0], [2, 2, 2, 2, 0, 0, 0, 0, 0, 0], [0] * 6 + [2] * 4, [0] * 5 + [2] * 5, [0] * 10]
    got = rg.object_segmentation_graphcut_pixels(segm, centres, gc_regul=.5, seed_size=1)
    assert got.tolist() == [[0] * 10, [1] * 5 + [0] * 5, [1] * 4 + [0] * 6, [0] * 6 + [2] * 4, [0] * 5 + [2] * 5, [0] * 10]


def test_object_shapes_golden():
    """region_growing.py:273-277, 309-319 (Ray features on the device, isb_ray_features_2d)"""
    rg = _rg()
    img = np.zeros((100, 100))
    img[20:70, 30:80] = 1
    rays, _ = rg.compute_segm_object_shape(img, ray_step=45)
    np.testing.assert_allclose(rays, [36.7, 26.0, 35.3, 25.0, 35.3, 25.0, 35.3, 26.0], atol=0.1)
    img1 = np.zeros((100, 100))
    img1[20:50, 30:60] = 1
    img1[40:80, 50:90] = 2
    img2 = np.zeros((100, 100))
    img2[10:40, 20:50] = 1
    img2[50:80, 20:50] = 1
    img2[50:80, 60:90] = 1
    list_rays, list_shifts = rg.compute_object_shapes([img1, img2], ray_step=45)
    assert np.array(list_rays).astype(int).tolist() == [[19, 17, 9, 17, 19, 14, 19, 14], [29, 21, 28, 20, 28, 20, 28, 21],
                                                       [22, 16, 21, 15, 21, 15, 21, 16], [22, 16, 21, 15, 21, 15, 21, 16],
                                                       [22, 16, 21, 15, 21, 15, 21, 16]]
    np.testing.assert_allclose(np.array(list_shifts) % 180, [135., 45., 45., 45., 45.], atol=1e-4)


def test_segm_prob_fg_golden():
    """region_growing.py:1144-1147"""
    slic = np.array([[0, 0, 0, 0, 1, 1, 1, 1], [2, 2, 2, 2, 3, 3, 3, 3]])
    segm = np.array([0, 1, 1, 0])[slic]
    np.testing.assert_allclose(_rg().compute_segm_prob_fg(slic, segm, [0.3, 0.8]), [0.3, 0.8, 0.8, 0.3])


def test_update_shape_costs_set_of_models_golden():
    """region_growing.py:896-934 (the mixture model only weights the two tables; the values below do not depend on its fit)"""
    from sklearn import mixture
    rg = _rg()
    np.random.seed(0)
    slic = np.kron(np.arange(16).reshape(4, 4), np.ones((2, 2), dtype=int))
    points = np.array([(y, x) for y in (0, 2, 4, 6) for x in (0, 2, 4, 6)])
    labels = np.array([0] * 4 + [0, 1, 1, 0, 0, 1, 1, 0] + [0] * 4)
    cdf1, cdf2 = np.zeros((8, 10)), np.zeros((8, 7))
    cdf1[:7] = 0.5
    cdf1[:4] = 1.0
    cdf2[:6] = 1.0
    set_m_cdf = [([4] * 8, cdf1), ([5] * 8, cdf2)]
    mm = mixture.GaussianMixture(2).fit(np.random.random((100, 8)))
    s_costs = np.zeros((len(points), 2))
    s_costs, centres, shifts, _ = rg.compute_update_shape_costs_points_close_mean_cdf(s_costs, slic, points, labels, [(0, 0)], [(np.inf, np.inf)],
                                                                                     [0], [0], (mm, set_m_cdf))
    assert centres.tolist() == [[3, 3]] and shifts.tolist() == [90.]
    got = np.round(s_costs, 3)
    assert np.all(got[:7, 1] == -0.01) and got[7, 1] == 0.868 and got[8, 1] == -0.01 and got[15, 1] == 4.605


def test_region_growing_graphcut_trajectories_golden():
    """region_growing.py:1517-1617: the criterion per growing step and the final label maps of the three doctest runs"""
    rg = _rg()
    h, w = 15, 20
    slic = _block_slic(h, w, 2)
    segm = _box(h, w, (3, 12), (5, 17))
    slic_prob_fg = rg.compute_segm_prob_fg(slic, segm, [0.1, 0.9])
    want = _box(h, w, (4, 12), (6, 16))
    dbg = {}
    labels = rg.region_growing_shape_slic_graphcut(slic, slic_prob_fg, [(7.5, 10)], (None, CHIST4), coef_pairwise=0, debug_history=dbg)
    assert np.round(dbg['criteria']).astype(int).tolist() == [397, 325, 206, 111, 81, 81]
    assert np.array_equal(labels[slic], want)
    labels = rg.region_growing_shape_slic_graphcut(slic, slic_prob_fg, [(7.5, 10)], (None, CHIST4), coef_pairwise=2, debug_history=dbg)
    assert np.round(dbg['criteria']).astype(int).tolist() == [415, 380, 289, 193, 164, 164]
    assert np.array_equal(labels[slic], want)
    segm = np.ones((h, w), dtype=int)
    chist = np.zeros((16, 9))
    chist[:, :5] = 1.
    slic_prob_fg = rg.compute_segm_prob_fg(slic, segm, [0.1, 0.9])
    dbg = {}
    labels = rg.region_growing_shape_slic_graphcut(slic, slic_prob_fg, [(6.5, 9)], (None, chist), coef_shape=10., coef_pairwise=1, debug_history=dbg)
    assert np.round(dbg['criteria']).astype(int).tolist() == [7506, 7120, 6328, 5719, 5719]
    want = _box(h, w, (4, 10), (4, 14)) | _box(h, w, (2, 12), (6, 12))
    assert np.array_equal(labels[slic], want)


def test_region_growing_greedy_trajectories_golden():
    """region_growing.py:1191-1291"""
    rg = _rg()
    h, w = 15, 20
    slic = _block_slic(h, w, 2)
    segm = _box(h, w, (3, 12), (5, 17))
    slic_prob_fg = rg.compute_segm_prob_fg(slic, segm, [0.1, 0.9])
    dbg = {}
    labels = rg.region_growing_shape_slic_greedy(slic, slic_prob_fg, [(7.5, 10)], (None, CHIST4), coef_pairwise=1, debug_history=dbg)
    crit = np.round(dbg['criteria']).astype(int).tolist()
    assert crit[:11] == [406, 352, 334, 316, 300, 283, 270, 254, 238, 226, 210] and crit[-2:] == [123, 123]
    assert np.array_equal(labels[slic], _box(h, w, (4, 12), (6, 16)))
    segm = np.ones((h, w), dtype=int)
    chist = np.zeros((16, 9))
    chist[:, :5] = 1.
    slic_prob_fg = rg.compute_segm_prob_fg(slic, segm, [0.1, 0.9])
    dbg = {}
    labels = rg.region_growing_shape_slic_greedy(slic, slic_prob_fg, [(6.5, 9)], (None, chist), coef_shape=10, coef_pairwise=1, debug_history=dbg)
    assert np.round(dbg['criteria']).astype(int).tolist() == [7506, 7120, 6715, 6328, 5719, 5719]
    assert np.array_equal(labels[slic], _box(h, w, (4, 10), (4, 14)) | _box(h, w, (2, 12), (6, 12)))


def test_rg2sp_recovers_synthetic_eggs_from_slic_superpixels():
    """config-4 shaped end-to-end run on synthetic data (the drosophila images of the reference do not travel to the GPU box):
    elliptic 'eggs' on a noisy background, SLIC superpixels from the device, a shape model from the eggs' own Ray features
    (cumulative histograms), RG2SP from the perturbed true centres -> every egg is recovered with a Jaccard index above 0.75"""
    import bench
    from pyimsegm_b200 import superpixels
    rg = _rg()
    img, annot, centres = bench.synth_eggs_image(11, 360, 480)
    slic = superpixels.segment_slic_img2d(img, sp_size=12, relative_compact=0.3)
    list_rays, _ = rg.compute_object_shapes([(annot == i + 1) for i in range(len(centres))], ray_step=10, interp_order='spline', smooth_coef=1)
    chist = rg.transform_rays_model_cdf_histograms(np.round(list_rays).astype(int), nb_bins=12)
    slic_prob_fg = rg.compute_segm_prob_fg(slic, (annot > 0).astype(int), [0.1, 0.9])
    start = [(c[0] + 4, c[1] - 5) for c in centres]
    labels = rg.region_growing_shape_slic_graphcut(slic, slic_prob_fg, start, (None, chist), 'cdf', coef_shape=2., coef_pairwise=5.,
                                                   prob_label_trans=[0.1, 0.03], nb_iter=50)
    segm = labels[slic]
    for i in range(len(centres)):
        a, b = segm == i + 1, annot == i + 1
        jac = (a & b).sum() / float((a | b).sum())
        assert jac > 0.75, 'egg %d: Jaccard %.2f' % (i, jac)
