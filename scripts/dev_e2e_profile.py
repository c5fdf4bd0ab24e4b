import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_image, FEATURES, SP_SIZE, SP_REGUL, NB_CLASSES
from pyimsegm_b200 import pipelines as pl
img = synth_image(2)
host = torch.from_numpy(img).pin_memory().numpy()
for _ in range(4):
    out = pl.pipe_color2d_slic_features_model_graphcut(host, NB_CLASSES, FEATURES, sp_size=SP_SIZE)
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    out = pl.pipe_color2d_slic_features_model_graphcut(host, NB_CLASSES, FEATURES, sp_size=SP_SIZE)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
