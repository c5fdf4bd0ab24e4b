"""dev: segment_images_batch throughput against the number of streams / images in flight (2048x2048 bench image)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_image, FEATURES, SP_SIZE, SP_REGUL
from pyimsegm_b200 import pipelines as pl
img = torch.from_numpy(synth_image(2)).pin_memory().numpy()
for ns, mif in ((2, 4), (3, 4), (3, 6), (4, 8), (1, 2)):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = pl.segment_images_batch([img] * 12, 3, FEATURES, sp_size=SP_SIZE, sp_regul=SP_REGUL, nb_streams=ns, max_in_flight=mif)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        del res
    print('streams', ns, 'in flight', mif, 'ms/image %.2f' % (dt / 12 * 1e3), 'MPix/s %.0f' % (12 * 4.194304 / dt))
