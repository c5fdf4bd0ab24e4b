"""small run of the tensor-core Leung-Malik path (whole image and row bands), the tcgen05 known-answer kernels and the SLIC sweeps for
compute-sanitizer --tool memcheck"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import synth_regions
from pyimsegm_b200 import _lib, pipelines as pl, tiled
from pyimsegm_b200.engine import get_engine
from pyimsegm_b200.superpixels import slic_params
torch = _lib.require_cuda()
lib = _lib.lib()
img, _ = synth_regions(100, 150, seed=2)
for key in ('tLM', 'tLM_short'):
    slic, fts = pl.compute_color2d_superpixels_features(img, {'color': ['mean'], key: ['mean', 'std', 'energy']}, sp_size=12, sp_regul=0.3)
    print(key, 'ok', fts.shape, bool(np.isfinite(fts).all()))
eng = get_engine()
n_seg, compact = slic_params(img.shape[:2], 12, 0.3)
res = tiled.slic_tiled(img, n_seg, compact, bands_per_rank=2, eng=eng, raw_margin=tiled.LM_ROW_MARGIN)
f = eng.to_host(tiled.texture_stats_tiled(res, img.dtype, ('mean', 'energy'), 'short', eng=eng))
print('banded texture ok', f.shape, bool(np.isfinite(f).all()))
rng = np.random.RandomState(0)
for variant, N, K in ((0, 80, 40), (2, 240, 64), (2, 48, 40)):
    A = torch.from_numpy(rng.standard_normal((128, K)).astype(np.float32)).cuda()
    B = torch.from_numpy(rng.standard_normal((N, K)).astype(np.float32)).cuda()
    D = torch.zeros((128, N), dtype=torch.float32, device='cuda')
    _lib.check(lib.isb_umma_selftest(_lib.ptr(A), _lib.ptr(B), N, K, variant, _lib.ptr(D), _lib.stream_ptr()))
    torch.cuda.synchronize()
print('umma ok')
