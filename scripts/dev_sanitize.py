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
