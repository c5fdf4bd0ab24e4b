#!/bin/bash
# dev: sweep warps per CTA x CTAs per SM of k_update on the GPU box (rebuilds the library per variant)
f=pyimsegm_b200/csrc/slic_kmeans.cu
for v in "8 4" "2 17" "2 18" "4 9" "1 32" "2 16"; do
  set -- $v
  sed -i "s/^constexpr int UWARPS = [0-9]*, UBLOCKS = [0-9]*;/constexpr int UWARPS = $1, UBLOCKS = $2;/" $f
  python -m pyimsegm_b200.build > /dev/null 2>&1 || { echo "build failed for $v"; continue; }
  cuobjdump -res-usage pyimsegm_b200/libimsegm_b200.so 2>/dev/null | grep -A1 "k_updateILb0" | grep -oE "REG:[0-9]+ STACK:[0-9]+"
  python -m pytest tests/test_gpu_parity.py -q -x -k "slic or label" 2>&1 | tail -1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('UWARPS=$1 UBLOCKS=$2', 'ms_per_step %.3f' % d['ms_per_step'], 'update %.3f' % d['stages']['slic_update']['ms_per_step'])"
done
