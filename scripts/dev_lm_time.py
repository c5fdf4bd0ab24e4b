import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_image
from pyimsegm_b200 import pipelines as pl, texture
from pyimsegm_b200.engine import get_engine
eng = get_engine()
H = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
img = synth_image(3, H, H)
res = pl._device_slic_features(eng, img, {'color': ['mean']}, 29, 0.2)
nb = int(eng.to_host(res.d_n_labels)[0])
for bank in ('normal', 'short'):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        texture.device_lm_features(eng, res.d_img, res.d_seg, nb, ('mean', 'std', 'energy'), bank)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nf = 76 if bank == 'normal' else 33
    print(bank, 'H', H, 'ms %.2f' % (dt * 1e3), 'algorithmic TFLOP/s %.1f' % (3 * nf * 33 * 33 * 2 * H * H / dt / 1e12))
lib = eng.lib
import ctypes as C
lib.isb_profile_enable(1)
texture.device_lm_features(eng, res.d_img, res.d_seg, nb, ('mean', 'std', 'energy'), 'normal')
