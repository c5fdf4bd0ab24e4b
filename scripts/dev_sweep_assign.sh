#!/bin/bash
# dev: sweep rows per thread / occupancy target of k_assign on the GPU box (rebuilds the library per variant)
f=pyimsegm_b200/csrc/slic_kmeans.cu
for v in "8 5" "4 3" "4 4" "4 5" "16 2" "16 3"; do
  set -- $v
  sed -i "s/^constexpr int AROWS = [0-9]*;/constexpr int AROWS = $1;/" $f
  sed -i "s/__global__ void __launch_bounds__(ATHREADS[^)]*) k_assign/__global__ void __launch_bounds__(ATHREADS, $2) k_assign/" $f
  python -m pyimsegm_b200.build > /dev/null 2>&1 || { echo "build failed for $v"; continue; }
  python -m pytest tests/test_gpu_parity.py -q -x -k "slic or label" 2>&1 | tail -1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('AROWS=$1 minBlocks=$2', 'ms_per_step %.3f' % d['ms_per_step'], 'assign %.3f' % d['stages']['slic_assign']['ms_per_step'])"
done
