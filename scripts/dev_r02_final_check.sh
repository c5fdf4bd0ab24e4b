#!/bin/bash
# last pass of the round on the final tree: the whole GPU test-suite, smoke, the bench line, config 3 / 4 lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python bench.py --workload config3 --steps 6 --warmup 2 > gpurun_out/r02_config3_n1.json 2> gpurun_out/r02_config3_n1.err; echo "config3 rc=$?"
python bench.py --workload config4 --steps 5 --warmup 3 > gpurun_out/r02_config4.json 2> gpurun_out/r02_config4.err; echo "config4 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print('value %.1f (%.3f ms) e2e %.1f (%.3f ms) pageable %.1f batch %.1f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_pageable']['value'], d['e2e_batch']['value'], d['gpu_launches']))
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','launch_ms')}, 'cpu', d['cpu_baseline']['value'], d.get('parity'), d['clocks'])
print({k: round(v['ms_per_step'],4) for k,v in d['stages'].items()})
for n in ('config3_n1','config4'):
    c=json.loads(open('gpurun_out/r02_%s.json' % n).read().strip().splitlines()[-1]); print(n, '%.1f %s %.2f ms' % (c['value'], c['unit'], c['ms_per_step']), {k: round(v['ms_per_step'],3) for k,v in c.get('stages',{}).items() if v['ms_per_step']>0.5}, c.get('ms_per_stage'))
PY
