#!/bin/bash
# dev: sweep the occupancy target of k_assign on the GPU box (rebuilds the library per variant)
for mb in 5 6 7 8; do
  sed -i "s/__global__ void __launch_bounds__(ATHREADS[^)]*) k_assign/__global__ void __launch_bounds__(ATHREADS, $mb) k_assign/" pyimsegm_b200/csrc/slic_kmeans.cu
  python -m pyimsegm_b200.build > /dev/null 2>&1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('minBlocks=$mb', 'ms_per_step %.3f' % d['ms_per_step'], 'assign %.3f' % d['stages']['slic_assign']['ms_per_step'])"
done
