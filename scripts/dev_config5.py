"""one-off: BASELINE config 5 size (8192x8192, ~80k superpixels) on ONE GPU, bit-exact against the oracle"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bench import synth_image
from pyimsegm_b200 import pipelines as pl
from sklearn import mixture, pipeline, preprocessing
H = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
img = synth_image(5, H, H, cell=128)
feats = {'color': ['mean']}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    slic, fts = pl.compute_color2d_superpixels_features(img, feats, sp_size=29, sp_regul=0.2)
    torch.cuda.synchronize(); print('gpu slic+features %.1f ms, N=%d' % ((time.perf_counter() - t0) * 1e3, slic.max() + 1))
t0 = time.perf_counter()
segm, soft = pl.pipe_color2d_slic_features_model_graphcut(img, 3, feats, sp_size=29, sp_regul=0.2)
print('gpu full pipeline (numpy in/out) %.1f ms' % ((time.perf_counter() - t0) * 1e3), np.bincount(segm.ravel()))
t0 = time.perf_counter()
slic_o, fts_o = oracle.compute_color2d_superpixels_features(img, ('mean',), 29, 0.2)
print('oracle slic+features %.1f s' % (time.perf_counter() - t0))
print('slic equal', np.array_equal(slic, slic_o), 'features close', np.allclose(fts, fts_o, rtol=1e-6, atol=1e-9))
model = pipeline.Pipeline([('s', preprocessing.StandardScaler()), ('m', mixture.GaussianMixture(3, random_state=0))]).fit(fts_o)
segm2, _ = pl.segment_color2d_slic_features_model_graphcut(img, model, feats, sp_size=29, sp_regul=0.2)
labels_o = oracle.segment_graph_cut_general(slic_o, model.predict_proba(fts_o), 1., 'model')
print('segmentation equal', np.array_equal(segm2, labels_o[slic_o]))
