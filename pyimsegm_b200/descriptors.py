"""
Per-superpixel descriptors on the GPU.

Mirror of the reference module ``imsegm/descriptors.py`` for the colour / texture statistics that the
SLIC -> features -> GraphCut pipeline uses (same public names, feature-dictionary grammar, column order and
error types).  The reference computes them in its only native module ``imsegm/features_cython.pyx``; here
they come from ``isb_segment_stats_2d`` (``include/imsegm_b200.h``).
"""
import itertools
import logging

import numpy as np

from .engine import FLAG_BITS, get_engine
from .utilities import ImageDimensionError

#: kept for API compatibility with the reference (descriptors.py:25-33); the native path here is CUDA and it
#: is always on -- there is no Python/NumPy fallback behind this switch
USE_CYTHON = True

#: all statistics computable on superpixels (reference descriptors.py:36)
NAMES_FEATURE_FLAGS = ('mean', 'std', 'energy', 'median', 'meanGrad')
#: sigmas of the Leung-Malik filter bank (reference descriptors.py:38-40)
DEFAULT_FILTERS_SIGMAS = (np.sqrt(2), 2, 2 * np.sqrt(2), 4)
SHORT_FILTERS_SIGMAS = (np.sqrt(2), 2, 4)
#: feature sets (reference descriptors.py:42-52)
FEATURES_SET_ALL = {
    'color': ('mean', 'std', 'energy', 'median', 'meanGrad'),
    'tLM': ('mean', 'std', 'energy', 'median', 'meanGrad'),
}
FEATURES_SET_COLOR = {'color': ('mean', 'std', 'energy')}
FEATURES_SET_TEXTURE = {'tLM': ('mean', 'std', 'energy')}
FEATURES_SET_TEXTURE_SHORT = {'tLM_short': ('mean', 'std', 'energy')}
HIST_CIRCLE_DIAGONALS = (10, 20, 30, 40, 50)
#: filter responses are clipped at this value (reference descriptors.py:55)
MAX_SIGNAL_RESPONSE = 1.e6


def _check_color_image_segm(image, segm):
    if image.shape[:2] != segm.shape:
        raise ImageDimensionError('ndarrays - image and segmentation do not match %r vs %r' % (image.shape, segm.shape))
    return True


def _check_gray_image_segm(image, segm):
    if image.shape != segm.shape:
        raise ImageDimensionError('ndarrays - image and segmentation do not match %r vs %r' % (image.shape, segm.shape))
    return True


def _check_color_image(image):
    if image.ndim != 3 or image.shape[2] != 3:
        raise ImageDimensionError('image is not RGB with dims %s' % repr(image.shape))
    return True


def _check_unrecognised_feature_group(feature_flags):
    unknown = [k for k in feature_flags if not (k.startswith('color') or k.startswith('tLM'))]
    if unknown:
        logging.warning('unrecognised following feature groups: %r', unknown)
    return unknown


def _check_unrecognised_feature_names(feature_flags):
    unknown = [k for k in feature_flags if k not in NAMES_FEATURE_FLAGS]
    if unknown:
        logging.warning('unrecognised following feature names: %r', unknown)
    return unknown


def _device_dtype(img):
    img = np.asarray(img)
    if img.dtype in (np.uint8, np.uint16, np.float32, np.float64):
        return img
    return img.astype(np.float64)


def _device_stats(img, seg, flags):
    """[nb, 3 * len(flags)] statistics in the order mean, std, energy (only those requested)"""
    img, seg = _device_dtype(img), np.asarray(seg)
    _check_color_image_segm(img, seg)
    eng = get_engine()
    nb = int(seg.max()) + 1
    d_img = eng.to_device(img, 'image')
    d_seg = eng.to_device(seg.astype(np.int32, copy=False), 'seg_in')
    feat, _, _ = eng.segment_stats(d_img, d_seg, nb, flags)
    return eng.to_host(feat).copy()


def cython_img2d_color_mean(img, seg):
    """ mean colour per segment, f32 pixels accumulated in f64 (reference descriptors.py:209-234) """
    return _device_stats(img, seg, ('mean', ))


def cython_img2d_color_energy(img, seg):
    """ mean squared colour per segment (reference descriptors.py:237-262) """
    return _device_stats(img, seg, ('energy', ))


def cython_img2d_color_std(img, seg, means=None):
    """ colour standard deviation per segment, two-pass about the f32 mean (reference descriptors.py:265-296).
    ``means`` is accepted for signature compatibility; the device path recomputes it in the same launch family """
    return _device_stats(img, seg, ('std', ))


def _host_label_sums(values, seg, nb):
    return np.stack([np.bincount(seg.ravel(), weights=values[..., c].ravel(), minlength=nb) for c in range(3)], axis=1)


def numpy_img2d_color_mean(img, seg):
    """ f64 host computation of the mean colour (the reference's NumPy variant, descriptors.py:299-332) """
    img, seg = np.asarray(img, dtype=float), np.asarray(seg)
    _check_color_image_segm(img, seg)
    nb = int(seg.max()) + 1
    cnt = np.bincount(seg.ravel(), minlength=nb).astype(float)
    cnt[cnt == 0] = -1
    return _host_label_sums(img, seg, nb) / cnt[:, None]


def numpy_img2d_color_std(img, seg, means=None):
    """ f64 host computation of the colour STD (reference descriptors.py:335-376) """
    img, seg = np.asarray(img, dtype=float), np.asarray(seg)
    _check_color_image_segm(img, seg)
    if means is None:
        means = numpy_img2d_color_mean(img, seg)
    nb = int(seg.max()) + 1
    if len(means) < nb:
        raise ValueError('number of means (%i) should be equal to number of labels (%i)' % (len(means), nb))
    cnt = np.bincount(seg.ravel(), minlength=nb).astype(float)
    cnt[cnt == 0] = -1
    var = _host_label_sums((img - np.asarray(means)[seg]) ** 2, seg, nb) / cnt[:, None]
    var[var == 0] = 0
    return np.sqrt(var)


def numpy_img2d_color_energy(img, seg):
    """ f64 host computation of the colour energy (reference descriptors.py:379-417) """
    img, seg = np.asarray(img, dtype=float), np.asarray(seg)
    _check_color_image_segm(img, seg)
    nb = int(seg.max()) + 1
    cnt = np.bincount(seg.ravel(), minlength=nb).astype(float)
    cnt[cnt == 0] = -1
    return _host_label_sums(img ** 2, seg, nb) / cnt[:, None]


def numpy_img2d_color_median(img, seg):
    """ per-segment, per-channel median (reference descriptors.py:420-455; no native path exists there either) """
    img, seg = np.asarray(img), np.asarray(seg)
    _check_color_image_segm(img, seg)
    nb = int(seg.max()) + 1
    flat = seg.ravel()
    order = np.argsort(flat, kind='stable')
    bounds = np.searchsorted(flat[order], np.arange(nb + 1))
    medians = np.full((nb, 3), np.nan)
    for c in range(3):
        vals = img[..., c].ravel()[order]
        for lb in range(nb):
            if bounds[lb + 1] > bounds[lb]:
                medians[lb, c] = np.median(vals[bounds[lb]:bounds[lb + 1]])
    return medians


def _device_gray_stats(img, seg, flags):
    import ctypes as C
    from . import _lib
    img, seg = _device_dtype(img), np.asarray(seg)
    _check_gray_image_segm(img, seg)
    eng = get_engine()
    nb = int(seg.max()) + 1
    d_img = eng.to_device(img, 'image_gray')
    d_seg = eng.to_device(seg.astype(np.int32, copy=False), 'seg_in')
    bits = sum(FLAG_BITS[f] for f in flags)
    feat = eng.buf('feat_gray', (nb, len(flags)), eng.torch.float64)
    wsb = eng.lib.isb_gray_stats_workspace_bytes(nb)
    ws = eng.buf('ws_gray', (wsb,), eng.torch.uint8)
    _lib.check(eng.lib.isb_gray_stats(_lib.ptr(d_img), _lib.DTYPE_CODES[str(img.dtype)], _lib.ptr(d_seg), C.c_longlong(img.size), nb, bits,
                                      _lib.ptr(feat), len(flags), 0, _lib.ptr(ws), C.c_size_t(wsb), _lib.stream_ptr()))
    return eng.to_host(feat).copy()


def cython_img3d_gray_mean(img, seg):
    """ mean intensity per segment of a gray volume (reference descriptors.py:458-487) """
    return _device_gray_stats(img, seg, ('mean', ))[:, 0]


def cython_img3d_gray_energy(img, seg):
    """ mean squared intensity per segment of a gray volume (reference descriptors.py:490-515) """
    return _device_gray_stats(img, seg, ('energy', ))[:, 0]


def cython_img3d_gray_std(img, seg, mean=None):
    """ intensity STD per segment of a gray volume, two-pass about the f32 mean (reference descriptors.py:518-551) """
    return _device_gray_stats(img, seg, ('std', ))[:, 0]


def cython_label_hist_seg2d(segm_select, struc_elem, nb_labels):
    """ histogram of the labels under a structuring element (reference descriptors.py:1479-1498) """
    import ctypes as C
    from . import _lib
    segm_select, struc_elem = np.array(segm_select, dtype=float), np.asarray(struc_elem)
    if segm_select.shape != struc_elem.shape:
        raise ValueError('segm. %r and mask %r sizes do not match' % (segm_select.shape, struc_elem.shape))
    segm_select[np.isnan(segm_select)] = -1
    eng = get_engine()
    d_a = eng.to_device(segm_select.astype(np.int16), 'hist_segm')
    d_b = eng.to_device(struc_elem.astype(np.int16), 'hist_selem')
    hist = eng.buf('hist_out', (int(nb_labels),), eng.torch.int32)
    _lib.check(eng.lib.isb_label_hist_2d(_lib.ptr(d_a), _lib.ptr(d_b), segm_select.shape[0], segm_select.shape[1], int(nb_labels),
                                         _lib.ptr(hist), _lib.stream_ptr()))
    return eng.to_host(hist).astype(float)


def ray_angle_tables(angle_step):
    """(sin, cos) float32 tables of the ray directions exactly as features_cython.pyx:247-268 forms them"""
    angles = np.arange(0, 360, angle_step, dtype=np.float32)
    rads = [float(np.float32(np.deg2rad(a))) for a in angles]
    return np.array([np.sin(r) for r in rads], dtype=np.float32), np.array([np.cos(r) for r in rads], dtype=np.float32)


def cython_ray_features_seg2d(seg_binary, position, angle_step=5., edge='up'):
    """ Ray features: distance from ``position`` to the first boundary along rays every ``angle_step`` degrees
    (reference descriptors.py:1628-1660).  ``position`` may also be an [n, 2] array: all positions run in one launch.

    :return ndarray: ray distances, float32 [n_angles] (or [n, n_angles])
    """
    import ctypes as C
    from . import _lib
    edge_int = {'down': -1, 'up': 1}[edge]
    seg = np.array(seg_binary, dtype=np.int8)
    pos = np.atleast_2d(np.array(position, dtype=np.int32))
    sin_a, cos_a = ray_angle_tables(float(angle_step))
    eng = get_engine()
    d_seg, d_pos = eng.to_device(seg, 'ray_seg'), eng.to_device(pos, 'ray_pos')
    d_s, d_c = eng.to_device(sin_a, 'ray_sin'), eng.to_device(cos_a, 'ray_cos')
    out = eng.buf('ray_out', (len(pos), len(sin_a)), eng.torch.float32)
    _lib.check(eng.lib.isb_ray_features_2d(_lib.ptr(d_seg), seg.shape[0], seg.shape[1], _lib.ptr(d_pos), len(pos), _lib.ptr(d_s), _lib.ptr(d_c),
                                           len(sin_a), edge_int, _lib.ptr(out), _lib.stream_ptr()))
    res = eng.to_host(out).copy()
    return res[0] if np.ndim(position) == 1 else res


def compute_image2d_color_statistic(image, segm, feature_flags=NAMES_FEATURE_FLAGS, color_name='color'):
    """ statistics of a colour image over the segments; columns are statistic-major, channel-minor
    (reference descriptors.py:787-863)

    :return tuple(ndarray,list(str)): features [nb_segments, 3 * nb_statistics], column names
    """
    image, segm = np.asarray(image), np.asarray(segm)
    _check_color_image(image)
    _check_color_image_segm(image, segm)
    ch_names = ['%s-ch%i' % (color_name, i + 1) for i in range(3)]
    native = [f for f in ('mean', 'std', 'energy') if f in feature_flags]
    blocks = {}
    if native:
        stats = _device_stats(image, segm, native)
        for i, f in enumerate(native):
            blocks[f] = stats[:, 3 * i:3 * i + 3]
    if 'median' in feature_flags:
        blocks['median'] = numpy_img2d_color_median(np.nan_to_num(image), segm)
    if 'meanGrad' in feature_flags:
        clean = np.nan_to_num(image)
        grad = np.zeros(clean.shape, dtype=clean.dtype if clean.dtype.kind == 'f' else float)
        for i in range(3):
            grad[:, :, i] = np.sum(np.gradient(clean[:, :, i]), axis=0)
        blocks['meanGrad'] = _device_stats(grad, segm, ('mean', ))
    order = [f for f in NAMES_FEATURE_FLAGS if f in feature_flags]
    nb = int(segm.max()) + 1
    features = np.hstack([blocks[f] for f in order]) if order else np.empty((nb, 0))
    names = list(itertools.chain.from_iterable(['%s_%s' % (n, f) for n in ch_names] for f in order))
    _check_unrecognised_feature_names(feature_flags)
    features = np.nan_to_num(features)
    features[features == 0] = 0
    if features.shape[1] != len(names):
        raise ValueError('features: %r and names %r' % (features.shape, names))
    return features, names


def norm_features(features, scaler=None):
    """ standardise the features (reference descriptors.py:866-877) """
    from sklearn import preprocessing
    if not scaler:
        scaler = preprocessing.StandardScaler()
        scaler.fit(features)
    return scaler.transform(features), scaler


def compute_selected_features_color2d(img, segments, feature_flags=FEATURES_SET_ALL):
    """ features of a colour image selected by the dictionary grammar ``{'color[_<space>]': flags, 'tLM[_short]': flags}``
    (reference descriptors.py:1207-1270)
    """
    img = np.asarray(img)
    _check_color_image(img)
    features, names = [], []
    for k in [k for k in feature_flags if k.startswith('color')]:
        clr = k.split('_')[-1] if '_' in k else 'rgb'
        if '_' in k:
            from .color import convert_img_color_from_rgb
            img_color = convert_img_color_from_rgb(img, clr)
        else:
            img_color = img
        fts, ns = compute_image2d_color_statistic(img_color, segments, feature_flags[k], color_name=clr)
        features.append(fts)
        names += ns
    for k in [k for k in feature_flags if k.startswith('tLM')]:
        bank_type = k.split('_')[-1] if '_' in k else 'normal'
        from .texture import compute_texture_desc_lm_img2d_clr
        fts, ns = compute_texture_desc_lm_img2d_clr(img, segments, feature_flags[k], bank_type)
        features.append(fts)
        names += ns
    _check_unrecognised_feature_group(feature_flags)
    features = np.concatenate(tuple(features), axis=1)
    features = np.nan_to_num(features)
    features[features == 0] = 0
    if not features.size:
        logging.error('not supported features: %r', feature_flags)
    if features.shape[1] != len(names):
        raise ValueError('features: %r and names %r' % (features.shape, names))
    return features, names


def compute_selected_features_gray2d(img, segments, features_flags=FEATURES_SET_ALL):
    """gray 2-D features go through the 3-D gray statistics in the reference (descriptors.py:1169-1204);
    that family is outside the accelerated hot path"""
    raise NotImplementedError('gray-image descriptors are outside the B200 hot path (SURVEY.md section 8f)')


def compute_selected_features_img2d(image, segm, features_flags=FEATURES_SET_COLOR):
    """ dispatch on the image kind (reference descriptors.py:1273-1285) """
    image = np.asarray(image)
    if image.ndim == 3 and image.shape[2] == 3:
        return compute_selected_features_color2d(image, segm, features_flags)
    if image.ndim == 2:
        return compute_selected_features_gray2d(image, segm, features_flags)
    logging.error('invalid image size - %r', image.shape)


def flags_are_native(dict_features):
    """True when every requested feature group / statistic is one the resident device path computes
    ('color', 'tLM', 'tLM_short' with mean / std / energy)"""
    return bool(dict_features) and all(k in ('color', 'tLM', 'tLM_short') and all(f in FLAG_BITS for f in v)
                                       for k, v in dict_features.items())


def native_feature_layout(dict_features):
    """[(key, flags, first column, n columns)] in the reference's column order: colour groups first, then texture"""
    layout, col = [], 0
    for k in [k for k in dict_features if k.startswith('color')] + [k for k in dict_features if k.startswith('tLM')]:
        flags = [f for f in ('mean', 'std', 'energy') if f in dict_features[k]]
        n = 3 * len(flags) * (1 if k == 'color' else (15 if k.endswith('_short') else 20))
        layout.append((k, flags, col, n))
        col += n
    return layout, col
