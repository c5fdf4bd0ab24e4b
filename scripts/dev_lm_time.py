"""dev tool: time the Leung-Malik stage on a 2048x2048 textured image (GPU) and print the stage timers"""
import ctypes as C
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyimsegm_b200 import _lib, pipelines
import torch

side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lib = _lib.lib()
img = bench.synth_texture_image(3000, side, side)
fts = {'color': ('mean', 'std', 'energy'), 'tLM': ('mean', 'std', 'energy')}
for _ in range(2):
    slic, f = pipelines.compute_color2d_superpixels_features(img, fts, sp_size=29, sp_regul=0.2)
lib.isb_profile_enable(1)
t0 = time.perf_counter()
n = 3
for _ in range(n):
    slic, f = pipelines.compute_color2d_superpixels_features(img, fts, sp_size=29, sp_regul=0.2)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n * 1e3
ns = lib.isb_profile_stage_count()
ms, cnt = (C.c_double * ns)(), (C.c_longlong * ns)()
lib.isb_profile_collect(ms, cnt)
lib.isb_profile_enable(0)
print('features %r, wall %.1f ms per call' % (f.shape, wall))
for i in range(ns):
    if ms[i] > 0:
        print('  %-18s %8.3f ms' % (lib.isb_profile_stage_name(i).decode(), ms[i] / n))
print('finite:', bool(np.isfinite(f).all()), ' |f| max %.4g' % np.abs(f).max())
