"""GPU parity of the Leung-Malik texture descriptors (imsegm/descriptors.py:1041-1106) against the float64 SciPy oracle.
Tolerance: the contraction runs as 3xTF32 on the tensor cores with FP32 accumulation (responses accurate to ~1e-6) and
the reference itself rounds every response to f32 before its statistics.  Every feature must agree within 2e-4 of the RMS
response of its battery and channel (see _check)."""
import numpy as np
import pytest

from conftest import synth_regions

pytestmark = pytest.mark.gpu


def _check(fts, want, flags=('mean', 'std', 'energy'), tol=2e-4):
    """errors are measured against the response scale of the battery/channel they belong to: rms = sqrt(max energy).
    (Edge / bar / Laplacian kernels are zero-mean: a segment's MEAN response is a small difference of large terms, so an
    error relative to the mean itself would measure cancellation, not the contraction.)"""
    nfl = len(flags)
    ie = flags.index('energy')
    worst = 0.0
    for b in range(fts.shape[1] // (3 * nfl)):
        for c in range(3):
            col = lambda i: b * 3 * nfl + i * 3 + c
            rms = np.sqrt(np.abs(want[:, col(ie)]).max()) + 1e-300
            for i, f in enumerate(flags):
                scale = rms * rms if f == 'energy' else rms
                worst = max(worst, np.abs(fts[:, col(i)] - want[:, col(i)]).max() / scale)
    assert worst < tol, 'max error relative to the battery response scale: %g' % worst
    return worst


@pytest.mark.parametrize('bank,shape', [('short', (30, 20)), ('normal', (70, 90)), ('short', (64, 48))])
def test_lm_descriptors_match_scipy_oracle(oracle, bank, shape):
    from oracle import texture as otex
    from pyimsegm_b200 import texture
    h, w = shape
    rng = np.random.RandomState(h)
    step = 5 if h == 30 else 11
    seg = (np.arange(h)[:, None] // step) * (-(-w // step)) + np.arange(w)[None, :] // step
    img = rng.random_sample((h, w, 3))
    img[: h // 2] += np.sin(np.arange(w) / 2.0)[None, :, None] * 0.3   # some oriented texture
    flags = ('mean', 'std', 'energy')
    fts, names = texture.compute_texture_desc_lm_img2d_clr(img, seg, flags, bank)
    want, want_names = otex.texture_desc_lm(img, seg, flags, bank)
    assert names == want_names and fts.shape == want.shape
    _check(fts, want)


def test_lm_reference_shape_contract():
    """descriptors.py:1052-1074: 24 segments x 135 features for the short bank with three statistics"""
    from pyimsegm_b200 import descriptors as ds
    from pyimsegm_b200 import texture
    np.random.seed(0)
    h, w, step = 30, 20, 5
    seg = (np.arange(h)[:, None] // step) * (w // step) + np.arange(w)[None, :] // step
    img = np.random.random((h, w, 3))
    fts, names = texture.compute_texture_desc_lm_img2d_clr(img, seg, ['mean', 'std', 'energy'], bank_type='short')
    assert fts.shape == (24, 135) and names[0] == 'tLM_sigma1.4-edge-ch1_mean' and names[-1] == 'tLM_sigma4.0-GaussLap2-ch3_energy'
    filters, fnames = texture.create_filter_bank_lm_2d(6, ds.SHORT_FILTERS_SIGMAS, 2)
    assert [f.shape for f in filters][:5] == [(2, 13, 13), (2, 13, 13), (1, 13, 13), (1, 13, 13), (1, 13, 13)]
    assert fnames[:2] == ['sigma1.4-edge', 'sigma1.4-bar']
    image = np.zeros((2, 10, 3))
    image[:, 2:6, 0] = 1
    segm = np.array([[0] * 5 + [1] * 5] * 2)
    features, _ = ds.compute_selected_features_color2d(image, segm, {'tLM_short': ('mean', 'energy')})
    assert features.shape == (2, 90)
    features, _ = ds.compute_selected_features_color2d(image, segm, {'color': ('mean', 'std', 'energy'), 'tLM': ('mean', 'std', 'energy')})
    assert features.shape == (2, 9 + 180)


def test_lm_on_slic_superpixels(oracle):
    from oracle import texture as otex
    from pyimsegm_b200 import texture
    img, _ = synth_regions(96, 128, seed=9)
    yy, xx = np.mgrid[:96, :128]
    img = np.clip(img + 0.2 * np.sin(xx / 3.0)[..., None] * (yy > 48)[..., None], 0, 1)
    seg = oracle.segment_slic_img2d(img, 16, 0.3)
    fts, _ = texture.compute_texture_desc_lm_img2d_clr(img, seg, ('mean', 'std', 'energy'), 'normal')
    want, _ = otex.texture_desc_lm(img, seg, ('mean', 'std', 'energy'), 'normal')
    _check(fts, want)
