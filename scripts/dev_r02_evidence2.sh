#!/bin/bash
# final evidence pass of round 2: the whole GPU test-suite, the bench line and the CPU arm, the other configurations, the ncu launch list
# of the bench command, full captures of the sweep kernels and of the Leung-Malik contraction, the tcgen05 issue-rate table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference arm rc=$?"
python bench.py --workload config3 --steps 4 --warmup 2 > gpurun_out/r02_config3_n1.json 2> gpurun_out/r02_config3_n1.err; echo "config3 rc=$?"
python bench.py --workload config4 --steps 3 --warmup 1 > gpurun_out/r02_config4.json 2> gpurun_out/r02_config4.err; echo "config4 rc=$?"
python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/r02_config5_n1.json 2> gpurun_out/r02_config5_n1.err; echo "config5 rc=$?"
timeout 300 python scripts/dev_umma_rate.py > gpurun_out/r02_umma_rate.txt 2>&1; echo "rate table rc=$?"
timeout 300 python scripts/dev_lm_time.py 2048 > gpurun_out/r02_lm_time.txt 2>&1; tail -3 gpurun_out/r02_lm_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_lm_launches.csv python scripts/dev_lm_time.py 2048 > /dev/null 2>&1; echo "lm launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lm_conv_ts -s 2 -c 1 -o gpurun_out/r02_lm_ts -f python scripts/dev_lm_time.py 2048 > /dev/null 2>&1; echo "ncu lm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assign -s 14 -c 1 -o gpurun_out/r02c_assign -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu assign rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update -s 14 -c 1 -o gpurun_out/r02c_update -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu update rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print('value %.1f (%.3f ms) e2e %.1f (%.3f ms) pageable %.1f batch %.1f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_pageable']['value'], d['e2e_batch']['value'], d['gpu_launches']))
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','launch_ms')}, 'cpu', d['cpu_baseline']['value'], d.get('parity'))
r=json.load(open('gpurun_out/r02_bench_reference.json'))
print('reference arm %.2f MPix/s, %.0f ms/step' % (r['value'], r['ms_per_step']))
for n in ('config3_n1','config4','config5_n1'):
    try:
        c=json.loads(open('gpurun_out/r02_%s.json' % n).read().strip().splitlines()[-1]); print(n, '%.1f %s %.2f ms' % (c['value'], c['unit'], c['ms_per_step']), {k: round(v['ms_per_step'],3) for k,v in c.get('stages',{}).items() if v['ms_per_step']>0.5})
    except Exception as e: print(n, 'no line', e)
PY
