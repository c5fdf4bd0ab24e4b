import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import oracle
from conftest import synth_regions
from pyimsegm_b200 import graph_cuts as gc, superpixels as sp
from pyimsegm_b200.engine import get_engine
img, _ = synth_regions(256, 320, seed=4)
seg = oracle.segment_slic_img2d(img, 16, 0.2)
rng = np.random.RandomState(1)
proba = rng.dirichlet(np.ones(3), seg.max() + 1)
c_g = np.array(sp.superpixel_centers(seg)); c_o = oracle.superpixel_centers(seg)
print('centres maxdiff', np.abs(c_g - c_o).max())
for et in ('', 'spatial', 'model', 'model_l1'):
    e_g, w_g = gc.compute_edge_weights(seg, proba=proba, edge_type=et)
    e_o, w_o = oracle.edge_weights(seg, proba, et)
    print(et, 'edges eq', np.array_equal(e_g, e_o), 'w maxrel', np.abs(w_g / w_o - 1).max())
edges = e_o
d = np.max((proba[edges[:, 0]] - proba[edges[:, 1]]) ** 2, axis=1)
print('std', np.std(d), 'mean', d.mean(), 'E', len(d))
w_ns = np.exp(-d / (2 * np.std(d) ** 2))
spd = oracle.spatial_dist(c_o, edges, True)
print('w_g*spd / w_ns', (w_g if False else gc.compute_edge_weights(seg, proba=proba, edge_type='model')[1] * spd / w_ns)[:8])
