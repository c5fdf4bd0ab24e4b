import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_image, FEATURES, SP_SIZE, SP_REGUL, NB_CLASSES
from pyimsegm_b200 import pipelines as pl
from pyimsegm_b200.engine import get_engine
eng = get_engine()
img = synth_image(2)
host = torch.from_numpy(img).pin_memory().numpy()
for _ in range(4):
    pl.pipe_color2d_slic_features_model_graphcut(host, NB_CLASSES, FEATURES, sp_size=SP_SIZE)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T()
    d = eng.to_device(host, 'image'); t1 = T()
    d_segm, d_soft, chk = pl._run_resident(eng, d, ('fit', NB_CLASSES, True, 99), FEATURES, SP_SIZE, SP_REGUL, 1.0, 'model'); t2e = time.perf_counter(); t2 = T()
    a = eng.pinned_empty(d_segm.shape, d_segm.dtype); b = eng.pinned_empty(d_soft.shape, d_soft.dtype); t3 = T()
    a.copy_(d_segm, non_blocking=True); b.copy_(d_soft, non_blocking=True); t4 = T()
    print('H2D %.2f | enqueue %.2f compute-total %.2f | pinned alloc %.2f | D2H %.2f ms' % ((t1-t0)*1e3, (t2e-t1)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3))
t0 = T()
for _ in range(5): out = pl.pipe_color2d_slic_features_model_graphcut(host, NB_CLASSES, FEATURES, sp_size=SP_SIZE)
print('e2e per call %.2f ms' % ((T()-t0)/5*1e3))
