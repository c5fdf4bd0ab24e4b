"""dev: time the large-D device class model (N x 189 features, 4 classes, 9 restarts) -- wall clock; run under ncu for the launch list"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyimsegm_b200 import graph_cuts as gc
from pyimsegm_b200.engine import get_engine
N, D, K = int(os.environ.get('GN', 4930)), 189, 4
rng = np.random.RandomState(1)
centers = rng.normal(0, 0.12, (K, D))
X = np.concatenate([c + rng.normal(0, 1.0, (N // K, D)) * rng.uniform(0.3, 1.0, D) for c in centers])
eng = get_engine()
d_feat = eng.to_device(X, 'feat_in')
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    proba, params = eng.gmm_fit_predict(d_feat, K, 9, 99, True, 0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    p = eng.to_host(params)
    tail = p[2 * D + K + K * D + 2 * K * D * D:]
    print('fit %.1f ms  lower %.4f n_iter %d conv %d best %d' % (dt * 1e3, tail[0], tail[1], tail[2], tail[4]))
