"""dev tool: banded Leung-Malik statistics against the whole-image descriptor, with diagnostics"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import synth_regions
from pyimsegm_b200.engine import get_engine
from pyimsegm_b200.superpixels import slic_params
from pyimsegm_b200.texture import device_lm_features
from pyimsegm_b200.tiled import LM_ROW_MARGIN, slic_tiled, texture_stats_tiled
eng = get_engine()
bank = sys.argv[1] if len(sys.argv) > 1 else 'normal'
rng = np.random.RandomState(5)
img = synth_regions(2000, 192, seed=21)[0] + 0.05 * rng.standard_normal((2000, 192, 3))
n_seg, compact = slic_params(img.shape[:2], 24, 0.2)
flags = ('mean', 'std', 'energy')
for nbands in (1, 3):
    res = slic_tiled(img, n_seg, compact, bands_per_rank=nbands, eng=eng, raw_margin=LM_ROW_MARGIN)
    print([repr(b) + ' up %d:%d' % (b.up_lo, b.up_hi) for b in res.bands])
    got = eng.to_host(texture_stats_tiled(res, img.dtype, flags, bank, eng=eng)).copy()
    seg = eng.to_host(res.d_seg).copy()
    d_img = eng.to_device(img, 'image')
    want = eng.to_host(device_lm_features(eng, d_img, res.d_seg, int(res.nb_bound), flags, bank)[0]).copy()
    d = np.abs(got - want)
    print('bands %d: max |diff| %.3g, max |want| %.3g, rel %.3g' % (nbands, d.max(), np.abs(want).max(), (d / (np.abs(want) + 1e-9)).max()))
    k, c = np.unravel_index(np.argmax(d), d.shape)
    ys = np.nonzero(seg == k)[0]
    print('  worst: segment %d (rows %d..%d) column %d got %.9g want %.9g' % (k, ys.min() if len(ys) else -1, ys.max() if len(ys) else -1, c, got[k, c], want[k, c]))
    bad = np.nonzero(d.max(axis=1) > 1e-7)[0]
    print('  segments off by > 1e-7: %d of %d' % (len(bad), int(seg.max()) + 1), [(int(b), int(np.nonzero(seg == b)[0].min())) for b in bad[:12]])
