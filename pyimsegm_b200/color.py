"""
Colour-space conversions for the ``color_<space>`` feature keys (reference ``imsegm/utilities/data_io.py:28-60`` ->
``skimage.color.rgb2hsv / rgb2luv / rgb2lab / rgb2hed / rgb2xyz``).  Host-side numpy restatement of the published
formulas (scikit-image is not a dependency here); the statistics over the converted image run on the device.
"""
import numpy as np

_XYZ_FROM_RGB = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
_D65 = np.array([0.95047, 1., 1.08883])
_RGB_FROM_HED = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11], [0.27, 0.57, 0.78]])


def _as_float(image):
    image = np.asarray(image)
    if image.dtype == np.uint8:
        return image / 255.
    if image.dtype == np.uint16:
        return image / 65535.
    return image.astype(float)


def rgb2xyz(rgb):
    arr = _as_float(rgb).copy()
    mask = arr > 0.04045
    arr[mask] = np.power((arr[mask] + 0.055) / 1.055, 2.4)
    arr[~mask] /= 12.92
    return arr @ _XYZ_FROM_RGB.T


def _xyz2lab(xyz):
    arr = xyz / _D65
    mask = arr > 0.008856
    arr[mask] = np.cbrt(arr[mask])
    arr[~mask] = 7.787 * arr[~mask] + 16. / 116.
    x, y, z = arr[..., 0], arr[..., 1], arr[..., 2]
    return np.stack([116. * y - 16., 500. * (x - y), 200. * (y - z)], axis=-1)


def rgb2lab(rgb):
    return _xyz2lab(rgb2xyz(rgb))


def rgb2luv(rgb):
    xyz = rgb2xyz(rgb)
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    eps = np.finfo(float).eps
    L = y / _D65[1]
    mask = L > 0.008856
    L[mask] = 116. * np.cbrt(L[mask]) - 16.
    L[~mask] = 903.3 * L[~mask]
    u0 = 4 * _D65[0] / np.dot([1, 15, 3], _D65)
    v0 = 9 * _D65[1] / np.dot([1, 15, 3], _D65)
    d = x + 15 * y + 3 * z + eps
    u = 13. * L * (4 * x / d - u0)
    v = 13. * L * (9 * y / d - v0)
    return np.stack([L, u, v], axis=-1)


def rgb2hsv(rgb):
    arr = _as_float(rgb)
    out = np.empty_like(arr)
    v = arr.max(-1)
    delta = np.ptp(arr, -1)
    old = np.seterr(invalid='ignore', divide='ignore')
    s = delta / v
    s[delta == 0.] = 0.
    idx = arr[..., 0] == v
    out[idx, 0] = (arr[idx, 1] - arr[idx, 2]) / delta[idx]
    idx = arr[..., 1] == v
    out[idx, 0] = 2. + (arr[idx, 2] - arr[idx, 0]) / delta[idx]
    idx = arr[..., 2] == v
    out[idx, 0] = 4. + (arr[idx, 0] - arr[idx, 1]) / delta[idx]
    h = (out[..., 0] / 6.) % 1.
    h[delta == 0.] = 0.
    np.seterr(**old)
    out[..., 0], out[..., 1], out[..., 2] = h, s, v
    out[np.isnan(out)] = 0
    return out


def rgb2hed(rgb):
    arr = np.maximum(_as_float(rgb), 1e-6)
    hed_from_rgb = np.linalg.inv(_RGB_FROM_HED)
    log_adjust = np.log(1e-6)
    stains = (np.log(arr) / log_adjust) @ hed_from_rgb
    return np.maximum(stains, 0)


#: conversion functions from RGB (reference data_io.py:28-34)
DICT_CONVERT_COLOR_FROM_RGB = {'hsv': rgb2hsv, 'luv': rgb2luv, 'lab': rgb2lab, 'hed': rgb2hed, 'xyz': rgb2xyz}


def convert_img_color_from_rgb(image, color_space):
    """ convert an RGB(A) image to ``color_space``; unknown spaces return the image unchanged (reference data_io.py:45-58) """
    image = np.asarray(image)
    if image.ndim == 3 and image.shape[-1] in (3, 4) and color_space in DICT_CONVERT_COLOR_FROM_RGB:
        image = DICT_CONVERT_COLOR_FROM_RGB[color_space](image[..., :3])
    return image
