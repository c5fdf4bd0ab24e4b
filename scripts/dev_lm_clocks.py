"""dev tool: SM clock and power while the Leung-Malik descriptor runs back to back for a few seconds (is the contraction power-capped?)"""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyimsegm_b200 import _lib
from pyimsegm_b200.engine import get_engine
from pyimsegm_b200.texture import device_lm_features
import torch
eng = get_engine()
img = bench.synth_texture_image(3000, 2048, 2048)
from pyimsegm_b200 import pipelines
slic, _ = pipelines.compute_color2d_superpixels_features(img, {'color': ('mean',)}, sp_size=29, sp_regul=0.2)
d_img = eng.to_device(img, 'image')
d_seg = eng.to_device(slic.astype(np.int32), 'seg_in')
nb = int(slic.max()) + 1
for _ in range(3):
    device_lm_features(eng, d_img, d_seg, nb, ('mean', 'std', 'energy'))
torch.cuda.synchronize()
out = open('gpurun_out/lm_clocks.csv', 'w')
smi = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,'
                        'clocks_event_reasons.sw_thermal_slowdown', '--format=csv,noheader,nounits', '-lms', '100'], stdout=out)
time.sleep(0.5)
n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
while time.perf_counter() - t0 < 4.0:
    for _ in range(10):
        device_lm_features(eng, d_img, d_seg, nb, ('mean', 'std', 'energy'))
    n += 10
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
smi.terminate(); smi.wait(); out.close()
rows = [l.strip().split(', ') for l in open('gpurun_out/lm_clocks.csv') if l.strip()]
rows = rows[6:-1]
clk = np.array([float(r[0]) for r in rows]); pw = np.array([float(r[2]) for r in rows])
print('%s: %d descriptors back to back, %.2f ms each; SM clock median %.0f MHz (min %.0f, max %.0f of %s), power median %.0f W (max %.0f); sw_power_cap active in %d of %d samples'
      % ('lm_texture', n, ms, np.median(clk), clk.min(), clk.max(), rows[0][1], np.median(pw), pw.max(),
         sum(r[3].strip() == 'Active' for r in rows), len(rows)))
