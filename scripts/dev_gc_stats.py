import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_image, FEATURES, SP_SIZE, SP_REGUL
from pyimsegm_b200 import pipelines as pl, graph_cuts as gc
from pyimsegm_b200.engine import get_engine
eng = get_engine()
img = synth_image(2)
res = pl._device_slic_features(eng, img, FEATURES, SP_SIZE, SP_REGUL)
nb = int(eng.to_host(res.d_n_labels)[0])
d_proba, _ = eng.gmm_fit_predict(res.d_feat[:nb], 3, 9, 99, True, 0)
pairwise = gc.compute_pairwise_cost(1.0, (nb, 3))
d_edges, d_n_edges, cap = eng.adjacency(res.d_seg, nb, 8 * nb)
E = int(eng.to_host(d_n_edges)[0])
print('N', nb, 'E', E)
_, _, unary_i, edge_wi, smooth_i = eng.gc_energies(d_proba, d_edges, E, None, res.d_centres, (1, 1), 1.0, pairwise)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    labels, energy, stats = eng.alpha_expansion(nb, 3, E, None, d_edges, edge_wi, unary_i, smooth_i, -1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('ms %.3f' % (dt * 1e3), 'energy', int(energy.item()), 'stats[moves,flows,sweeps,relabels,levels,smem]', stats.cpu().numpy().tolist())
print(np.bincount(labels.cpu().numpy()))
