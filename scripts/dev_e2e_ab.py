"""dev: single-image e2e with and without the early segm_soft download, alternating on the same box"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_image, FEATURES, SP_SIZE, SP_REGUL
from pyimsegm_b200 import pipelines as pl
img = torch.from_numpy(synth_image(2)).pin_memory().numpy()
def run(flag, n=10):
    pl.EARLY_SOFT_DOWNLOAD = flag
    keep = None
    for _ in range(3):
        keep = pl.pipe_color2d_slic_features_model_graphcut(img, 3, FEATURES, sp_size=SP_SIZE, sp_regul=SP_REGUL)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        keep = pl.pipe_color2d_slic_features_model_graphcut(img, 3, FEATURES, sp_size=SP_SIZE, sp_regul=SP_REGUL)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, keep
ref = None
for rep in range(3):
    for flag in (False, True):
        ms, out = run(flag)
        if ref is None:
            ref = out
        same = bool(np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1]))
        print('early_soft', flag, 'ms/image %.3f' % ms, 'identical results', same)
