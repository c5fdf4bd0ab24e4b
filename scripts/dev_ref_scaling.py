import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import cpu_reference_throughput
for w in (1, 8, 16, 32):
    v, dt, _ = cpu_reference_throughput(w, w)
    print('workers', w, 'images', w, 'wall %.1f s' % dt, '%.2f MPix/s' % v, flush=True)
