#!/bin/bash
# dev: cluster-parallel class model + graphs: whole gpu suite, headline bench, config 5 on one GPU with stage timers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests2.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/r02_gpu_tests2.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench3.json'))
print('value %.1f MPix/s (%.3f ms)  e2e %.1f (%.3f ms) pageable %.1f (%.3f ms) batch %.1f  launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_pageable']['value'], d['e2e_pageable']['ms_per_step'], d['e2e_batch']['value'], d['gpu_launches']))
print(d.get('stages_note'))
for k,v in d['stages'].items(): print('  %-18s %.4f ms (%g launches)' % (k, v['ms_per_step'], v['launches_per_step']))
PY
tail -3 gpurun_out/r02_bench3.err
timeout 600 python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/r02_config5_n1.json 2> gpurun_out/r02_config5_n1.err; echo "config5 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_config5_n1.json'))
print('config5 N=1: %.1f MPix/s (%.2f ms); whole-image API %.2f ms' % (d['value'], d['ms_per_step'], d['whole_image_single_gpu']['ms_per_step']))
for k,v in d['stages'].items(): print('  %-18s %.4f ms (%g launches)' % (k, v['ms_per_step'], v['launches_per_step']))
PY
tail -3 gpurun_out/r02_config5_n1.err
