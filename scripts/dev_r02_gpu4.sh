#!/bin/bash
# dev: parity of the SLIC path after a kernel change + the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiled.py tests/test_gpu_round2.py -q -x > gpurun_out/r02_parity.log 2>&1; echo "parity rc=$?"; tail -8 gpurun_out/r02_parity.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench2.json'))
print('value %.1f MPix/s (%.3f ms)  e2e %.1f (%.3f ms) pageable %.1f (%.3f ms) batch %.1f  launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_pageable']['value'], d['e2e_pageable']['ms_per_step'], d['e2e_batch']['value'], d['gpu_launches']))
print(d.get('stages_note'))
for k,v in d['stages'].items(): print('  %-18s %.4f ms (%g launches)' % (k, v['ms_per_step'], v['launches_per_step']))
PY
tail -3 gpurun_out/r02_bench2.err
