#!/bin/bash
# evidence pass of round 2: the bench line, the CPU arm, the ncu launch list of the bench command, full captures of the sweep kernels
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_assign -s 14 -c 1 -o gpurun_out/r02c_assign -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu assign rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_update -s 14 -c 1 -o gpurun_out/r02c_update -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu update rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print('value %.1f (%.3f ms) e2e %.1f (%.3f ms) pageable %.1f batch %.1f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_pageable']['value'], d['e2e_batch']['value'], d['gpu_launches']))
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','launch_ms')}, 'cpu', d['cpu_baseline']['value'], d['parity'])
r=json.load(open('gpurun_out/r02_bench_reference.json'))
print('reference arm %.2f MPix/s, %.0f ms/step, ideal %.0f ms' % (r['value'], r['ms_per_step'], r['cpu_baseline']['ideal_ms_per_step']))
PY
