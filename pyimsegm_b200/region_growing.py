"""
Region growing with a shape prior on superpixels (RG2SP) and object segmentation by GraphCut -- mirror of the reference module
``imsegm/region_growing.py`` (same names, arguments, return values; BASELINE config 4).

What runs where: everything that is sized by pixels or by the superpixel graph is a call into the CUDA library -- the region /
annotation histogram (``isb_region_label_hist``), the superpixel graph and centres (``isb_adjacency_edges``,
``isb_segment_stats_2d``), the Ray features of an object (``isb_ray_features_2d``) and every max-flow: the reference calls
``gco.cut_general_graph`` once per growing step (region_growing.py:1698,1715) and ``gco.cut_grid_graph`` for the pixel-level
variant (:248); here both are the device alpha-expansion (``graph_cuts.cut_general_graph`` / ``cut_grid_graph``).  The growing loop
itself and the shape-model bookkeeping (a few thousand superpixels, a handful of objects) are vectorised numpy on the host: the
shape prior of ALL superpixels about an object is one table lookup with bilinear interpolation instead of one scipy ``interp2d``
object per superpixel.
"""
import logging

import numpy as np
from scipy import ndimage, stats

from .descriptors import compute_ray_features_segm_2d, interpolate_ray_dist, shift_ray_features
from .graph_cuts import MAX_PAIRWISE_COST, compute_spatial_dist, cut_general_graph, cut_grid_graph, get_vertexes_edges
from .labeling import histogram_regions_labels_norm
from .superpixels import get_neighboring_segments, make_graph_segm_connect_grid2d_conn4, superpixel_centers

#: all infinity values in Graph-Cut terms are replaced by this value (reference region_growing.py:27)
GC_REPLACE_INF = 1e5
#: minimal value of any shape-prior probability (reference :29)
MIN_SHAPE_PROB = 0.01
#: maximal value of a unary (being a class) probability in Graph-Cut (reference :31)
MAX_UNARY_PROB = 1 - 0.01
#: thresholds of the iterative region growing (reference :33-38)
RG2SP_THRESHOLDS = {
    'centre': 30,  # min centre displacement since the last shape-prior update
    'shift': 15,  # min rotation change since the last update
    'volume': 0.1,  # min relative volume change since the last update
    'centre_init': 50,  # maximal move from the initial estimate
}


def _round_int(values):
    return np.round(values).astype(int)


def _labels_of_superpixels(slic, segm):
    return np.argmax(histogram_regions_labels_norm(slic, segm), axis=1)


def _distance_prior(dist, shape_mean, shape_std):
    """1 - normal CDF of the (integer part of the) distance: the radial shape prior of the two object_segmentation_* functions"""
    cdf = stats.norm.cdf(np.arange(int(np.max(dist) + 1)), shape_mean, shape_std)
    return (1. - cdf + 1e-9)[dist.astype(int)]


def object_segmentation_graphcut_slic(slic, segm, centres, labels_fg_prob=(0.1, 0.9), gc_regul=1, edge_coef=0.5, edge_type='model',
                                      coef_shape=0., shape_mean_std=(50., 10.), add_neighbours=False, debug_visual=None):
    """ object segmentation by one Graph Cut on the superpixel graph: label 0 = background, label i = the object about ``centres[i-1]``
    (reference region_growing.py:42-156)

    :param ndarray slic: superpixel map
    :param ndarray segm: input (class) segmentation
    :param list(tuple(int,int)) centres: one point per object
    :param list(float) labels_fg_prob: probability of each segmentation label to be foreground
    :return ndarray: label per superpixel, int32
    """
    labels_fg_prob = np.asarray(labels_fg_prob, dtype=float)
    if np.min(labels_fg_prob) >= 1:
        raise ValueError('non label can ce strictly 1')
    slic, segm = np.asarray(slic), np.asarray(segm)
    labels = _labels_of_superpixels(slic, segm)
    if segm.max() > len(labels_fg_prob):
        raise ValueError('table of label prob is shorter then the nb of labels in segmentation')
    if not list(centres):
        raise ValueError('at least one center has to be given')
    centres = [_round_int(c) for c in centres]
    nb_obj = len(centres)
    fg, bg = labels_fg_prob[labels], 1. - labels_fg_prob[labels]
    slic_points = np.asarray(superpixel_centers(slic))

    proba = np.column_stack([bg] + [fg] * nb_obj)
    shape = np.ones_like(proba)
    if coef_shape > 0:
        shape[:, 0] = bg
        for i, centre in enumerate(centres):
            dist = np.sqrt(np.sum((slic_points - centre) ** 2, axis=1))
            shape[:, i + 1] = _distance_prior(dist, *shape_mean_std)
    _, edges = get_vertexes_edges(slic)
    edges = np.array(edges)

    unary_cost = -np.log(proba) - coef_shape * np.log(shape)
    for i, pos in enumerate(centres):
        vertex = slic[tuple(pos)]
        unary_cost[vertex, i + 1] = 0
        if add_neighbours:      # the seed superpixel's neighbours are seeds too; their edges are taken out of the graph
            touching = np.logical_or(edges[:, 0] == vertex, edges[:, 1] == vertex)
            unary_cost[np.unique(edges[touching]), i + 1] = 0
            edges[touching] = 0
    unary_cost = np.maximum(unary_cost, -np.log(MAX_UNARY_PROB))

    if edge_type == 'model':
        dist = np.abs(fg[edges[:, 0]] - fg[edges[:, 1]])
        edge_weights = np.exp(-dist / (2 * np.std(dist) ** 2)) / compute_spatial_dist(slic_points, edges, relative=True)
    else:
        edge_weights = np.ones(len(edges))
    edge_weights = edge_weights * edge_coef
    pairwise_cost = (1 - np.eye(nb_obj + 1)) * gc_regul

    logging.debug('perform GraphCut')
    graph_labels = cut_general_graph(edges, edge_weights, unary_cost, pairwise_cost, n_iter=999)
    if debug_visual is not None:
        debug_visual['unary_imgs'] = [unary_cost[:, i][slic] for i in range(unary_cost.shape[-1])]
    return graph_labels


def _disk(radius):
    yy, xx = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    return (yy ** 2 + xx ** 2) <= radius ** 2


def object_segmentation_graphcut_pixels(segm, centres, labels_fg_prob=(0.1, 0.9), gc_regul=1, seed_size=0, coef_shape=0.,
                                        shape_mean_std=(50., 10.), debug_visual=None):
    """ object segmentation by one Graph Cut on the 4-connected PIXEL grid (reference region_growing.py:159-256)

    :param ndarray segm: input (class) segmentation
    :param list(tuple(int,int)) centres: one point per object
    :param int seed_size: radius of the disc about each centre that is forced to the object
    :return ndarray: label per pixel, int32 [H, W]
    """
    labels_fg_prob = np.asarray(labels_fg_prob, dtype=float)
    if np.min(labels_fg_prob) >= 1:
        raise ValueError('non label can ce strictly 1')
    segm = np.asarray(segm)
    if segm.max() > len(labels_fg_prob):
        raise ValueError('table of label proba is shorter then the nb of labels in segmentation')
    if not list(centres):
        raise ValueError('at least one center has to be given')
    height, width = segm.shape
    centres = [_round_int(c) for c in centres]
    nb_obj = len(centres)
    fg, bg = labels_fg_prob[segm], 1. - labels_fg_prob[segm]

    proba = np.stack([bg] + [fg] * nb_obj, axis=-1)
    shape = np.ones_like(proba)
    if coef_shape > 0:
        shape[:, :, 0] = bg
        rows, cols = np.mgrid[:height, :width]
        for i, centre in enumerate(centres):
            dist = np.sqrt((rows - centre[0]) ** 2 + (cols - centre[1]) ** 2)
            shape[:, :, i + 1] = _distance_prior(dist, *shape_mean_std)
    unary = -np.log(proba) - coef_shape * np.log(shape)
    for i, pos in enumerate(centres):
        if seed_size > 0:
            mask = np.zeros(segm.shape, dtype=bool)
            r0, c0 = pos[0] - seed_size, pos[1] - seed_size
            selem = _disk(seed_size)
            rs, cs = slice(max(r0, 0), min(r0 + selem.shape[0], height)), slice(max(c0, 0), min(c0 + selem.shape[1], width))
            mask[rs, cs] = selem[rs.start - r0:rs.stop - r0, cs.start - c0:cs.stop - c0]
            unary[np.logical_and(mask, segm > 0), i + 1] = 0
        else:
            unary[pos[0], pos[1], i + 1] = 0
    pairwise = (1 - np.eye(nb_obj + 1)) * gc_regul
    cost_v, cost_h = np.ones((height - 1, width)), np.ones((height, width - 1))
    segm_obj = cut_grid_graph(unary, pairwise, cost_v, cost_h, n_iter=999).reshape(segm.shape)
    if debug_visual is not None:
        debug_visual['unary_imgs'] = [unary[:, :, i] for i in range(unary.shape[-1])]
    return segm_obj


# ---------------------------------------------------------------------------------------------------------------------
# shape models from Ray features (reference region_growing.py:259-588)
# ---------------------------------------------------------------------------------------------------------------------

def compute_segm_object_shape(img_object, ray_step=5, interp_order=3, smooth_coef=0, shift_method='phase'):
    """ Ray features of ONE object about its centre of gravity, missing rays interpolated, optionally smoothed, rotated to the
    dominant direction (reference region_growing.py:259-286)

    :return tuple(list(float), float): ray distances, rotation shift in degrees
    """
    img_object = np.asarray(img_object)
    centre = [int(round(c)) for c in ndimage.center_of_mass(img_object)]
    ray_dist = compute_ray_features_segm_2d(img_object, centre, ray_step, 0, edge='down')
    if interp_order is not None and -1 in ray_dist:
        ray_dist = interpolate_ray_dist(ray_dist, interp_order)
    if smooth_coef > 0:
        ray_dist = ndimage.gaussian_filter1d(ray_dist, smooth_coef)
    ray_dist, shift = shift_ray_features(ray_dist, shift_method)
    return ray_dist.tolist(), shift


def compute_object_shapes(list_img_objects, ray_step=5, interp_order=3, smooth_coef=0, shift_method='phase'):
    """ :func:`compute_segm_object_shape` of every object of every image; an image with at most one non-zero label is split into its
    connected components first (reference region_growing.py:289-331)

    :return tuple(list(list(float)), list(float)):
    """
    list_rays, list_shifts = [], []
    for img_objects in list_img_objects:
        img_objects = np.asarray(img_objects)
        uq_labels = np.unique(img_objects)
        if len(uq_labels) <= 2:
            img_objects, _ = ndimage.label(img_objects)
            uq_labels = np.unique(img_objects)
        for label in uq_labels[1:]:
            rays, shift = compute_segm_object_shape(img_objects == label, ray_step, interp_order, smooth_coef, shift_method)
            list_rays.append(rays)
            list_shifts.append(shift)
    return list_rays, list_shifts


def compute_cumulative_distrib(means, stds, weights, max_dist):
    """ inverse cumulative distribution of a mixture of normal distributions per ray direction, sampled at integer distances and
    min-max normalised (reference region_growing.py:334-361)

    :param ndarray means: [nb_components, nb_rays]
    :param ndarray stds: [nb_components, nb_rays]
    :param ndarray weights: [nb_components]
    :return ndarray: [nb_rays, int(max_dist) + 1]
    """
    means, stds = np.asarray(means, dtype=float), np.asarray(stds, dtype=float)
    samples = np.arange(int(max_dist) + 1)
    cdist = []
    for i in range(means.shape[1]):
        cdf = np.zeros(int(max_dist + 1))
        for j, w in enumerate(weights):
            cdf += stats.norm.cdf(samples, means[j, i], stds[j, i]) * w
        cdf = (cdf - cdf.min()) / (cdf.max() - cdf.min())
        cdist.append(1. - cdf + 1e-9)
    return np.array(cdist)


def _diag_stds(mm):
    covs = mm.covariances if hasattr(mm, 'covariances') else mm.covariances_
    return np.sqrt(abs(covs))[:, np.eye(mm.means_.shape[1], dtype=bool)]


def transform_rays_model_cdf_mixture(list_rays, coef_components=1):
    """ Bayesian Gaussian mixture over the ray vectors (number of components from MeanShift) turned into one cumulative
    distribution per ray (reference region_growing.py:364-401)

    :return tuple(model, list(list(float))):
    """
    from sklearn import cluster, mixture
    rays = np.array(list_rays)
    ms = cluster.MeanShift().fit(rays)
    nb_components = int(len(np.unique(ms.labels_)) * coef_components)
    mm = mixture.BayesianGaussianMixture(n_components=nb_components)
    mm.fit(rays, ms.labels_)
    max_dist = np.max([[m[i] + np.sqrt(c[i, i]) for i in range(len(m))] for m, c in zip(mm.means_, mm.covariances_)])
    cdist = compute_cumulative_distrib(mm.means_, _diag_stds(mm), mm.weights_, max_dist)
    return mm, cdist.tolist()


def transform_rays_model_sets_mean_cdf_mixture(list_rays, nb_components=5, slic_size=15):
    """ diagonal Bayesian Gaussian mixture; every component becomes its own (smoothed mean, cumulative distribution) pair
    (reference region_growing.py:404-438)

    :return tuple(model, list(tuple(list(float), ndarray))):
    """
    from sklearn import mixture
    rays = np.array(list_rays)
    mm = mixture.BayesianGaussianMixture(n_components=nb_components, covariance_type='diag')
    mm.fit(rays)
    list_mean_cdf = []
    for mean, covar in zip(mm.means_, mm.covariances_):
        std = ndimage.gaussian_filter1d(np.sqrt(covar + 1) * 2 + slic_size, 1)
        mean = ndimage.gaussian_filter1d(mean, 1)
        cdist = compute_cumulative_distrib(np.array([mean]), np.array([std]), np.array([1]), np.max(mean + 2 * std))
        list_mean_cdf.append((mean.tolist(), cdist))
    return mm, list_mean_cdf


def transform_rays_model_sets_mean_cdf_kmeans(list_rays, nb_components=5):
    """ k-means over the ray vectors; every cluster becomes its own (smoothed mean, cumulative distribution) pair
    (reference region_growing.py:441-470) """
    from sklearn import cluster
    rays = np.array(list_rays)
    kmeans = cluster.KMeans(nb_components).fit(rays)
    list_mean_cdf = []
    for lb, mean in enumerate(kmeans.cluster_centers_):
        std = ndimage.gaussian_filter1d(np.std(rays[kmeans.labels_ == lb], axis=0), 1)
        mean = ndimage.gaussian_filter1d(mean, 1)
        std = (std + 1) * 5.
        cdist = compute_cumulative_distrib(np.array([mean]), np.array([std]), np.array([1]), np.max(mean + 2 * std))
        list_mean_cdf.append((mean.tolist(), cdist))
    return kmeans, list_mean_cdf


def _cdist_of_clusters(rays, labels, means, smooth_means=False):
    stds = np.zeros((len(means), rays.shape[1]))
    for i, lb in enumerate(np.unique(labels)):
        if smooth_means:
            means[i, :] = ndimage.gaussian_filter1d(np.mean(rays[labels == lb], axis=0), 1)
        stds[i, :] = np.std(rays[labels == lb], axis=0)
    stds += 1
    weights = np.bincount(labels) / float(len(labels))
    max_dist = np.max(means + stds)
    return compute_cumulative_distrib(means, stds, weights, max_dist)


def transform_rays_model_cdf_spectral(list_rays, nb_components=5):
    """ spectral clustering of the ray vectors -> mixture of per-cluster normal distributions -> cumulative distribution
    (reference region_growing.py:473-510) """
    from sklearn import cluster
    rays = np.array(list_rays)
    sc = cluster.SpectralClustering(nb_components).fit(rays)
    means = np.zeros((len(np.unique(sc.labels_)), rays.shape[1]))
    cdist = _cdist_of_clusters(rays, sc.labels_, means, smooth_means=True)
    return sc, cdist.tolist()


def transform_rays_model_cdf_kmeans(list_rays, nb_components=None):
    """ k-means (number of clusters from MeanShift when not given) -> cumulative distribution (reference region_growing.py:513-554) """
    from sklearn import cluster
    rays = np.array(list_rays)
    if not nb_components:
        ms = cluster.MeanShift().fit(rays)
        kmeans = cluster.KMeans(len(np.unique(ms.labels_)))
        kmeans.fit(rays, ms.labels_)
    else:
        kmeans = cluster.KMeans(nb_components).fit(rays)
    cdist = _cdist_of_clusters(rays, kmeans.labels_, kmeans.cluster_centers_)
    return kmeans, cdist.tolist()


def transform_rays_model_cdf_histograms(list_rays, nb_bins=10):
    """ cumulative histogram of the measured distances per ray direction (reference region_growing.py:557-588)

    :return list(list(float)):
    """
    rays = np.array(list_rays)
    max_dist = np.max(rays)
    list_chist = []
    for i in range(rays.shape[1]):
        cum = np.zeros(max_dist + 1)
        hist, bin_edges = np.histogram(rays[:, i], nb_bins)
        hist = hist.astype(float) / np.sum(hist)
        bin_edges = bin_edges.astype(int)
        bins = ((bin_edges[1:] + bin_edges[:-1]) / 2).astype(int)
        cum[:bins[0]] = 1
        for j, edge in enumerate(bins):
            cum[edge:] = cum[edge - 1] - hist[j]
        list_chist.append(cum.tolist())
    return list_chist


# ---------------------------------------------------------------------------------------------------------------------
# shape prior of points about an object (reference region_growing.py:591-1062)
# ---------------------------------------------------------------------------------------------------------------------

def shape_prior_table_cdf_points(points, cum_distribution, centre, angle_shift=0):
    """ :func:`compute_shape_prior_table_cdf` of MANY points in one pass: angle (measured like the Ray features, rotated by
    ``angle_shift``) and distance of every point from ``centre`` index the table [nb_angles, nb_distances]; bilinear interpolation
    inside the table, the last column of the nearest angle beyond it

    :param ndarray points: [N, 2]
    :return ndarray: [N]
    """
    table = np.asarray(cum_distribution, dtype=float)
    points = np.asarray(points, dtype=float).reshape(-1, 2)
    angle_step = 360. / table.shape[0]
    table = np.vstack((table, table[:1]))           # the angle axis is periodic
    dx, dy = points[:, 0] - centre[0], points[:, 1] - centre[1]
    dist = np.sqrt(dx ** 2 + dy ** 2)
    angle = ((2 * 360) + 90 - np.rad2deg(np.arctan2(dy, dx)) - angle_shift) % 360
    angle_norm = angle / angle_step
    prior = np.empty(len(points))
    far = dist >= (table.shape[1] - 1)
    prior[far] = table[np.rint(angle_norm[far]).astype(int), -1]
    near = ~far
    a0, d0 = np.floor(angle_norm[near]).astype(int), np.floor(dist[near]).astype(int)
    if np.any(a0 >= table.shape[0] - 1):
        raise ValueError('angle %i is larger then size %i' % (a0.max(), table.shape[0]))
    fa, fd = angle_norm[near] - a0, dist[near] - d0
    prior[near] = (table[a0, d0] * (1 - fa) * (1 - fd) + table[a0 + 1, d0] * fa * (1 - fd)
                   + table[a0, d0 + 1] * (1 - fa) * fd + table[a0 + 1, d0 + 1] * fa * fd)
    return prior


def compute_shape_prior_table_cdf(point, cum_distribution, centre, angle_shift=0):
    """ shape prior of one point from the centre, the rotation and the cumulative-distribution table (reference
    region_growing.py:591-649)

    :return float:
    """
    return float(shape_prior_table_cdf_points([point], cum_distribution, centre, angle_shift)[0])


def compute_centre_moment_points(points):
    """ centre of a point set and the direction (degrees) of its principal axis (reference region_growing.py:704-747)

    :return tuple(ndarray, float):
    """
    points = np.asarray(points, dtype=float)
    centre = np.mean(points, axis=0)
    theta = 0
    if len(points) > 1:
        evals, evecs = np.linalg.eig(np.cov((points - centre).T))
        axis = evecs[:, np.argmax(evals)]
        theta = np.arctan2(axis[0], axis[1])
    theta = (360 + round(np.rad2deg(theta))) % 360
    return centre, float(theta)


def _clamp_to_init(centre_new, init_centre, max_move):
    """a centre farther than ``max_move`` from the initial estimate is pulled back onto that circle"""
    move2 = np.sum((np.asarray(centre_new) - np.asarray(init_centre)) ** 2)
    if move2 > max_move ** 2:
        return init_centre + (max_move / np.sqrt(move2)) * (np.asarray(centre_new) - np.asarray(init_centre))
    return centre_new


def compute_update_shape_costs_points_table_cdf(lut_shape_cost, points, labels, init_centres, centres, shifts, volumes, shape_chist,
                                                selected_idx=None, swap_shift=False, dict_thresholds=None):
    """ update the shape cost of every object whose centre or orientation moved more than the thresholds since its prior was last
    evaluated; one cumulative-distribution table for all objects (reference region_growing.py:750-852)

    :return tuple: lut_shape_cost, centres, shifts, volumes
    """
    if len(points) != len(labels):
        raise ValueError('number of points (%i) and labels (%i) should match' % (len(points), len(labels)))
    points, labels = np.asarray(points), np.asarray(labels)
    selected = np.arange(len(points)) if selected_idx is None else np.asarray(list(selected_idx), dtype=int)
    thresholds = RG2SP_THRESHOLDS if dict_thresholds is None else dict_thresholds
    _, cdf = shape_chist
    for i, centre in enumerate(centres):
        centre_new, shift = compute_centre_moment_points(points[labels == i + 1])
        centre_new = _round_int(centre_new)
        if swap_shift:
            shift = (shift + 90) % 360
            shifts[i] = shift
        centre_new = _clamp_to_init(centre_new, init_centres[i], thresholds['centre_init'])
        moved2 = np.sum((np.array(centre_new) - np.array(centre)) ** 2)
        centre_moved = moved2 > thresholds['centre'] ** 2
        shift_moved = np.abs(shift - shifts[i]) > thresholds['shift']
        if not centre_moved and not shift_moved and not swap_shift:
            continue
        if centre_moved:
            centres[i] = np.asarray(centre_new).tolist()
        if shift_moved:
            shifts[i] = shift
        shape_proba = np.zeros(len(points))
        shape_proba[selected] = shape_prior_table_cdf_points(points[selected], cdf, centres[i], shifts[i])
        lut_shape_cost[:, i + 1] = -np.log(shape_proba + MIN_SHAPE_PROB)
    lut_shape_cost[np.isinf(lut_shape_cost)] = GC_REPLACE_INF
    return lut_shape_cost, np.array(centres), np.array(shifts, dtype=float), volumes


def compute_update_shape_costs_points_close_mean_cdf(lut_shape_cost, slic, points, labels, init_centres, centres, shifts, volumes,
                                                     shape_model_cdfs, selected_idx=None, swap_shift=False, dict_thresholds=None):
    """ as :func:`compute_update_shape_costs_points_table_cdf` with a SET of shape models: the Ray features of the object's current
    segmentation weight the models' tables through the mixture model's ``predict_proba`` (reference region_growing.py:855-990)

    :return tuple: lut_shape_cost, centres, shifts, volumes
    """
    if len(points) != len(labels):
        raise ValueError('number of points (%i) and labels (%i) should match' % (len(points), len(labels)))
    points, labels = np.asarray(points), np.asarray(labels)
    selected = np.arange(len(points)) if selected_idx is None else np.asarray(list(selected_idx), dtype=int)
    thresholds = RG2SP_THRESHOLDS if dict_thresholds is None else dict_thresholds
    segm_obj = labels[slic]
    model, list_mean_cdf = shape_model_cdfs
    list_cdfs = [np.asarray(cdf) for _, cdf in list_mean_cdf]
    angle_step = 360 / len(list_cdfs[0])
    for i, centre in enumerate(centres):
        centre_new, shift = compute_centre_moment_points(points[labels == i + 1])
        centre_new = _round_int(centre_new)
        rays, _ = compute_segm_object_shape(segm_obj == i + 1, angle_step, smooth_coef=0)
        if swap_shift:
            shift = (shift + 90) % 360
            shifts[i] = shift
        volume = np.sum(labels == (i + 1))
        volume_diff = 0 if volumes[i] == 0 else np.abs(volume - volumes[i]) / float(volumes[i])
        centre_new = _clamp_to_init(centre_new, init_centres[i], thresholds['centre_init'])
        moved2 = np.sum((np.array(centre_new) - np.array(centre)) ** 2)
        centre_moved = moved2 > thresholds['centre'] ** 2
        shift_moved = np.abs(shift - shifts[i]) > thresholds['shift']
        volume_moved = volume_diff > thresholds['volume']
        if not (centre_moved or shift_moved or volume_moved or swap_shift):
            continue
        if centre_moved:
            centres[i] = np.asarray(centre_new).tolist()
        if shift_moved:
            shifts[i] = shift
        if volume_moved:
            volumes[i] = volume
        weights = model.predict_proba([rays]).ravel()
        cdist = np.zeros(np.max([cdf.shape for cdf in list_cdfs], axis=0))
        for j, cdf in enumerate(list_cdfs):
            cdist[:, :cdf.shape[1]] += weights[j] * cdf
        shape_proba = np.zeros(len(points))
        shape_proba[selected] = shape_prior_table_cdf_points(points[selected], cdist, centres[i], shifts[i])
        lut_shape_cost[:, i + 1] = -np.log(shape_proba + MIN_SHAPE_PROB)
    lut_shape_cost[np.isinf(lut_shape_cost)] = GC_REPLACE_INF
    return lut_shape_cost, np.array(centres), np.array(shifts, dtype=float), volumes


def compute_data_costs_points(slic, slic_prob_fg, centres, labels):
    """ look-up table of the data cost per superpixel and label; the superpixel under every centre is given to its object
    (reference region_growing.py:993-1011)

    :return tuple(ndarray, ndarray): lut_data_cost [N, nb_objects + 1], labels
    """
    slic_prob_fg = np.asarray(slic_prob_fg, dtype=float)
    data_proba = np.column_stack([1. - slic_prob_fg] + [slic_prob_fg] * len(centres))
    for i, centre in enumerate(centres):
        labels[slic[centre[0], centre[1]]] = i + 1
    lut_data_cost = -np.log(data_proba + 1e-9)
    lut_data_cost[np.isinf(lut_data_cost)] = GC_REPLACE_INF
    return lut_data_cost, labels


def update_shape_costs_points(lut_shape_cost, slic, points, labels, init_centres, centres, shifts, volumes, shape_model, shape_type,
                              selected_idx=None, swap_shift=False, dict_thresholds=None):
    """ dispatch on the kind of shape model: 'cdf' (one table) or 'set_cdfs' (a mixture of tables) (reference
    region_growing.py:1014-1062) """
    thresholds = RG2SP_THRESHOLDS if dict_thresholds is None else dict_thresholds
    if shape_type == 'cdf':
        return compute_update_shape_costs_points_table_cdf(lut_shape_cost, points, labels, init_centres, centres, shifts, volumes,
                                                           shape_model, selected_idx, swap_shift, thresholds)
    if shape_type == 'set_cdfs':
        return compute_update_shape_costs_points_close_mean_cdf(lut_shape_cost, slic, points, labels, init_centres, centres, shifts,
                                                                volumes, shape_model, selected_idx, swap_shift, thresholds)
    raise NameError('Not supported type of shape model "%s"' % shape_type)


def compute_pairwise_penalty(edges, labels, prob_bg_fg=0.05, prob_fg1_fg2=0.01):
    """ cost of every edge whose end points carry different labels: -log of the background/object or object/object transition
    probability (reference region_growing.py:1065-1085) """
    lb = np.asarray(labels)[np.asarray(edges)]
    differs = lb[:, 0] != lb[:, 1]
    touches_bg = np.logical_and(differs, np.logical_or(lb[:, 0] == 0, lb[:, 1] == 0))
    costs = -np.log(prob_fg1_fg2) * differs
    costs[touches_bg] = -np.log(prob_bg_fg)
    return costs


def get_neighboring_candidates(slic_neighbours, labels, object_idx, use_other_obj=True):
    """ superpixels adjacent to object ``object_idx`` that could join it: background ones, and with ``use_other_obj`` those of
    other objects (reference region_growing.py:1088-1111)

    :return list(int):
    """
    labels = np.asarray(labels)
    members = np.flatnonzero(labels == object_idx)
    near = np.unique(np.fromiter((n for m in members for n in slic_neighbours[m]), dtype=np.int64))
    if use_other_obj:
        return [int(n) for n in near if labels[n] != object_idx]
    return [int(n) for n in near if labels[n] == 0]


def compute_rg_crit(labels, lut_data_cost, lut_shape_cost, slic_weights, edges, coef_data, coef_shape, coef_pairwise, prob_label_trans):
    """ the energy the region growing minimises: size-weighted data + shape costs of the labelling plus the transition penalties
    (reference region_growing.py:1114-1133) """
    idx = np.arange(len(labels))
    crit = np.sum(slic_weights * (coef_data * lut_data_cost[idx, labels] + coef_shape * lut_shape_cost[idx, labels]))
    if coef_pairwise > 0:
        pairwise_costs = compute_pairwise_penalty(edges, labels, prob_label_trans[0], prob_label_trans[1])
        pairwise_costs[np.isinf(pairwise_costs)] = GC_REPLACE_INF
        crit += coef_pairwise * np.sum(pairwise_costs)
    return crit


def compute_segm_prob_fg(slic, segm, labels_prob):
    """ foreground probability of every superpixel = the table entry of its majority label (reference region_growing.py:1136-1152) """
    return np.array(labels_prob)[_labels_of_superpixels(np.asarray(slic), np.asarray(segm))]


# ---------------------------------------------------------------------------------------------------------------------
# region growing (reference region_growing.py:1155-1729)
# ---------------------------------------------------------------------------------------------------------------------

class _GrowingState(object):
    """what both growing strategies share: the superpixel graph, the cost tables and the per-object shape state"""

    def __init__(self, slic, slic_prob_fg, centres, shape_model, shape_type, dict_thresholds, bg_offset):
        slic_prob_fg = np.asarray(slic_prob_fg, dtype=float)
        if len(slic_prob_fg) < np.max(slic):
            raise ValueError('dims of probs %s and slic %s not match' % (len(slic_prob_fg), np.max(slic)))
        self.slic = slic
        self.thresholds = RG2SP_THRESHOLDS if dict_thresholds is None else dict_thresholds
        self.points = _round_int(superpixel_centers(slic))
        self.weights = np.bincount(slic.ravel())
        self.init_centres = _round_int(centres)
        _, self.edges = make_graph_segm_connect_grid2d_conn4(slic)
        self.neighbours = get_neighboring_segments(self.edges)
        self.csr = _neighbours_csr(self.neighbours)
        self.shape_model, self.shape_type = shape_model, shape_type
        labels = np.zeros(len(self.points), dtype=int)
        self.lut_data_cost, self.labels = compute_data_costs_points(slic, slic_prob_fg, self.init_centres, labels)
        self.lut_shape_cost = np.empty((len(labels), len(self.init_centres) + 1))
        self.lut_shape_cost[:, 0] = -np.log(1 - slic_prob_fg + bg_offset)
        self.centres = np.ones(np.asarray(self.init_centres).shape) * np.inf
        self.shifts = np.zeros(len(self.init_centres))
        self.volumes = [1] * len(self.shifts)
        self.update_shape(False)

    def update_shape(self, swap_shift):
        self.lut_shape_cost, self.centres, self.shifts, self.volumes = update_shape_costs_points(
            self.lut_shape_cost, self.slic, self.points, self.labels, self.init_centres, self.centres, self.shifts, self.volumes,
            self.shape_model, self.shape_type, None, swap_shift, self.thresholds)

    def crit(self, labels, coefs):
        return compute_rg_crit(labels, self.lut_data_cost, self.lut_shape_cost, self.weights, self.edges, *coefs)

    def record(self, debug_history, crit):
        if debug_history is not None:
            debug_history['labels'].append(self.labels.copy())
            debug_history['criteria'].append(crit)
            debug_history['centres'].append(self.centres.copy())
            debug_history['shifts'].append(self.shifts.tolist())
            debug_history['lut_shape_cost'].append(self.lut_shape_cost.copy())


def _init_history(debug_history, state):
    if debug_history is not None:
        debug_history.update({'criteria': [], 'labels': [], 'centres': [], 'shifts': [], 'lut_data_cost': state.lut_data_cost.copy(),
                              'lut_shape_cost': []})


def region_growing_shape_slic_greedy(slic, slic_prob_fg, centres, shape_model, shape_type='cdf', coef_data=1., coef_shape=1, coef_pairwise=1,
                                     prob_label_trans=(.1, .01), allow_obj_swap=True, greedy_tol=1e-3, dict_thresholds=None, nb_iter=999,
                                     debug_history=None):
    """ region growing with a shape prior on superpixels, greedy strategy: per step every neighbouring superpixel whose move to the
    adjacent object lowers the energy by (almost) the best amount is moved (reference region_growing.py:1155-1388)

    :return ndarray: object label per superpixel
    """
    slic = np.asarray(slic)
    state = _GrowingState(slic, slic_prob_fg, centres, shape_model, shape_type, dict_thresholds, bg_offset=0.)
    coefs = (coef_data, coef_shape, coef_pairwise, prob_label_trans)
    list_swap_shift = [False]
    _init_history(debug_history, state)
    for _ in range(nb_iter):
        state.labels = enforce_center_labels(slic, state.labels, state.centres)
        state.record(debug_history, state.crit(state.labels, coefs))
        candidates, objs_idx = [], []
        for i in range(len(state.centres)):
            near = get_neighboring_candidates(state.neighbours, state.labels, i + 1, allow_obj_swap)
            candidates += near
            objs_idx += [i + 1] * len(near)
        state.update_shape(list_swap_shift[-1])
        crit = state.crit(state.labels, coefs)
        scores = []
        for idx, lb in zip(objs_idx, candidates):
            labels_new = state.labels.copy()
            labels_new[lb] = idx
            scores.append((idx, lb, crit - state.crit(labels_new, coefs)))
        scores = sorted(scores, key=lambda x: x[2], reverse=True)
        if not scores or scores[0][2] < 0:
            if any(list_swap_shift[-7:]):      # the orientation was already shaken recently: nothing left to gain
                break
            list_swap_shift.append(True)
        else:
            list_swap_shift.append(False)
        best_score = scores[0][2]
        for obj, sp, score in scores:
            if (best_score - score) / best_score < greedy_tol and score > 0:
                state.labels[sp] = obj
    return state.labels


def _neighbours_csr(slic_neighbours):
    """list of neighbour lists -> (indptr, indices) arrays"""
    counts = np.fromiter((len(n) for n in slic_neighbours), dtype=np.int64, count=len(slic_neighbours))
    indptr = np.concatenate(([0], np.cumsum(counts)))
    indices = np.fromiter((v for n in slic_neighbours for v in n), dtype=np.int64, count=int(indptr[-1]))
    return indptr, indices


def prepare_graphcut_variables(candidates, slic_points, slic_neighbours, slic_weights, labels, nb_centres, lut_data_cost, lut_shape_cost,
                               coef_data, coef_shape, coef_pairwise, prob_label_trans, _csr=None):
    """ the sub-graph of one growing step: the candidate superpixels (free, but only towards labels present among their neighbours)
    plus their non-candidate neighbours (pinned to their current label) (reference region_growing.py:1391-1464; the reference grows
    the vertex list and the unary table one neighbour at a time, here the whole band is assembled with array operations)

    :return tuple: vertexes (superpixel indices), edges [E, 2] (indices into vertexes), edge_weights, unary, pairwise
    """
    slic_points, labels = np.asarray(slic_points), np.asarray(labels)
    cand = np.asarray(candidates, dtype=np.int64)
    if np.max(cand) >= len(slic_points):
        raise ValueError('max candidate idx: %d for %d centres' % (np.max(cand), len(slic_points)))
    indptr, indices = _neighbours_csr(slic_neighbours) if _csr is None else _csr
    if indices.size and indices.max() >= len(slic_points):
        raise ValueError('max slic neighbours idx: %d for %d centres' % (indices.max(), len(slic_points)))
    nb_labels = nb_centres + 1
    nc = len(cand)
    # every (candidate slot, neighbour) pair, candidate by candidate, neighbours in list order
    counts = indptr[cand + 1] - indptr[cand]
    slot = np.repeat(np.arange(nc), counts)
    offset = np.arange(int(counts.sum())) - np.repeat(np.cumsum(counts) - counts, counts)
    nbr = indices[indptr[cand][slot] + offset]
    # free vertices: the data + shape cost of the labels that occur among the neighbours, "infinite" otherwise
    cost = np.asarray(slic_weights)[cand][:, None] * (coef_data * lut_data_cost[cand] + coef_shape * lut_shape_cost[cand])
    present = np.zeros((nc, nb_labels), dtype=bool)
    present[slot, labels[nbr]] = True
    unary_free = np.where(present, cost, GC_REPLACE_INF)
    # vertex numbering: candidates in list order (a repeated one keeps its first slot), then new neighbours as first met
    position = np.full(len(slic_points), -1, dtype=np.int64)
    position[cand[::-1]] = np.arange(nc)[::-1]
    fresh = nbr[position[nbr] < 0]
    _, first = np.unique(fresh, return_index=True)
    pinned = fresh[np.sort(first)]
    position[pinned] = nc + np.arange(len(pinned))
    vertexes = cand.tolist() + pinned.tolist()
    unary_pinned = np.full((len(pinned), nb_labels), GC_REPLACE_INF)
    unary_pinned[np.arange(len(pinned)), labels[pinned]] = 0
    unary = np.maximum(np.vstack([unary_free, unary_pinned]), -np.log(MAX_UNARY_PROB))
    edges = np.stack([slot, position[nbr]], axis=1)
    edge_weights = np.ones(len(edges)) / compute_spatial_dist(slic_points[vertexes], edges, relative=True)
    pairwise = np.full((nb_labels, nb_labels), -np.log(prob_label_trans[0]))
    pairwise[1:, 1:] = -np.log(prob_label_trans[1])
    np.fill_diagonal(pairwise, 0)
    pairwise = np.minimum(pairwise * coef_pairwise, MAX_PAIRWISE_COST)
    return vertexes, edges, edge_weights, unary, pairwise


def enforce_center_labels(slic, labels, centres):
    """ the superpixel under every object's centre always carries that object's label (reference region_growing.py:1467-1479) """
    for i, center in enumerate(centres):
        labels[slic[int(center[0]), int(center[1])]] = i + 1
    return labels


def region_growing_shape_slic_graphcut(slic, slic_prob_fg, centres, shape_model, shape_type='cdf', coef_data=1., coef_shape=1,
                                       coef_pairwise=2, prob_label_trans=(0.1, 0.03), optim_global=True, allow_obj_swap=True,
                                       dict_thresholds=None, nb_iter=999, debug_history=None):
    """ region growing with a shape prior on superpixels (RG2SP): per step one Graph Cut over the band of superpixels around the
    current objects decides which of them join, leave or change object (reference region_growing.py:1482-1729).  Every cut is the
    device alpha-expansion.

    :param ndarray slic: superpixel map
    :param list(float) slic_prob_fg: foreground probability per superpixel
    :param list(tuple(int,int)) centres: initial centre per object
    :param shape_model: (model, table) for ``shape_type`` 'cdf' or (mixture model, [(mean, table), ...]) for 'set_cdfs'
    :param bool optim_global: one cut for all objects per step (True) or one cut per object
    :return ndarray: object label per superpixel
    """
    slic = np.asarray(slic)
    state = _GrowingState(slic, slic_prob_fg, centres, shape_model, shape_type, dict_thresholds, bg_offset=1e-9)
    coefs = (coef_data, coef_shape, coef_pairwise, prob_label_trans)
    labels_history = [np.zeros(len(state.points), dtype=int)]
    list_swap_shift = [False]
    _init_history(debug_history, state)

    def cut_band(candidates, labels_gc):
        gc_vertexes, gc_edges, edge_weights, unary, pairwise = prepare_graphcut_variables(
            candidates, state.points, state.neighbours, state.weights, state.labels, len(state.centres), state.lut_data_cost,
            state.lut_shape_cost, coef_data, coef_shape, coef_pairwise, prob_label_trans, _csr=state.csr)
        if len(gc_edges) > 0:
            labels_gc[gc_vertexes] = cut_general_graph(gc_edges, edge_weights, unary, pairwise, n_iter=999)

    for _ in range(nb_iter):
        state.labels = enforce_center_labels(slic, state.labels, state.centres)
        state.record(debug_history, state.crit(state.labels, coefs))
        labels_gc = state.labels.copy()
        if optim_global:
            candidates = []
            for i in range(len(state.centres)):
                candidates += get_neighboring_candidates(state.neighbours, state.labels, i + 1, allow_obj_swap)
            state.update_shape(list_swap_shift[-1])
            cut_band(candidates, labels_gc)
        else:
            for i in range(len(state.centres)):
                candidates = get_neighboring_candidates(state.neighbours, state.labels, i + 1, allow_obj_swap)
                state.update_shape(list_swap_shift[-1])
                cut_band(candidates, labels_gc)
        if np.array_equal(state.labels, labels_gc):
            # a fixed point: shake the orientation once, stop when that was already tried or the labelling has been seen before
            seen = any(np.array_equal(labels_gc, old) for old in labels_history[:-1])
            if any(list_swap_shift[-2:]) or seen:
                break
            list_swap_shift.append(True)
        else:
            list_swap_shift.append(False)
        state.labels = labels_gc
        labels_history.append(state.labels.copy())
    return state.labels
