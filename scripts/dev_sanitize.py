"""small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import synth_regions
from pyimsegm_b200 import pipelines as pl, graph_cuts as gc, texture
img, _ = synth_regions(96, 136, seed=1)
segm, soft = pl.pipe_color2d_slic_features_model_graphcut(img, 3, {'color': ['mean', 'std', 'energy']}, sp_size=12, sp_regul=0.3)
print('pipe ok', segm.shape, np.bincount(segm.ravel()))
slic, fts = pl.compute_color2d_superpixels_features(img, {'color': ['mean'], 'tLM_short': ['mean', 'energy']}, sp_size=12, sp_regul=0.3)
print('lm ok', fts.shape)
rng = np.random.RandomState(0)
n = 600; a = rng.randint(0, n, 2000); b = rng.randint(0, n, 2000); k = a != b
e = np.unique(np.stack([np.minimum(a, b)[k], np.maximum(a, b)[k]], 1), axis=0).astype(np.int32)
lab = gc.cut_general_graph(e, rng.rand(len(e)) + 0.01, gc.compute_unary_cost(rng.dirichlet(np.ones(4) * 2, n)), gc.compute_pairwise_cost(1.5, (n, 4)))
print('gc ok', np.bincount(lab))
# row-band mode (three bands in one process) and the large-D class model
from pyimsegm_b200 import tiled
from pyimsegm_b200.superpixels import slic_params
n_seg, compact = slic_params(img.shape[:2], 12, 0.3)
res = tiled.slic_tiled(img, n_seg, compact, bands_per_rank=3)
print('banded slic ok', res.fell_back)
res = tiled.slic_tiled(img, n_seg, compact, bands_per_rank=2, slic_zero=True)
print('banded slico ok', res.fell_back)
s2, p2, rows = tiled.pipe_color2d_slic_features_model_graphcut_tiled(img, 3, {'color': ['mean', 'std']}, sp_size=12, sp_regul=0.3, bands_per_rank=2)
print('banded pipe ok', s2.shape, rows, bool(np.array_equal(s2, pl.pipe_color2d_slic_features_model_graphcut(img, 3, {'color': ['mean', 'std']}, sp_size=12, sp_regul=0.3)[0])))
X = np.concatenate([c + rng.normal(0, 0.5, (120, 24)) for c in rng.normal(0, 2.0, (3, 24))])
m = gc.estim_class_model(X, 3)
print('large-D gmm ok', np.bincount(m.predict_proba(X).argmax(1)))
# gray-volume path and the descriptor drivers
from pyimsegm_b200 import descriptors as ds, superpixels as sp
vol = rng.random_sample((6, 40, 44))
vol[:, :, 22:] += 1.0
sv = sp.segment_slic_img3d_gray(vol, 8, 0.3, (2, 1, 1))
print('volume slic ok', sv.max() + 1)
print('volume pipe ok', np.bincount(pl.pipe_gray3d_slic_features_model_graphcut(vol, 2, {'color': ['mean', 'std']}, spacing=(2, 1, 1), sp_size=8).ravel()))
print('volume texture ok', ds.compute_texture_desc_lm_img3d_val(vol[:2, :20, :24], sv[:2, :20, :24], ('mean', 'energy'), 'short')[0].shape)
pts = [(0, 0), (95, 135), (40, 60)]
print('hist ok', ds.compute_label_histograms_positions(segm, pts, [2, 5, 9])[0].shape, ds.compute_label_hist_segm(segm, (3, 4), np.ones((7, 5)), 3)[1])
print('ray ok', ds.compute_ray_features_positions(segm, pts, 30, border_labels=[0])[0].shape)
