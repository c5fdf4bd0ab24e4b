#!/bin/bash
# dev: remaining tests, config 4 / config 3 lines, fresh source-level captures of the sweep kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_region_growing.py tests/test_gpu_round2.py tests/test_gpu_descriptor_drivers.py -q > gpurun_out/r02_rg_tests2.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02_rg_tests2.log
timeout 600 python bench.py --workload config4 --steps 3 --warmup 2 > gpurun_out/r02_config4.json 2> gpurun_out/r02_config4.err; echo "config4 rc=$?"; tail -c 1300 gpurun_out/r02_config4.json; tail -5 gpurun_out/r02_config4.err
timeout 600 python bench.py --workload config3 --steps 5 --warmup 2 > gpurun_out/r02_config3_n1.json 2> gpurun_out/r02_config3_n1.err; echo "config3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_config3_n1.json'))
print('config3: %.1f ms/image (%.1f MPix/s), features only %.1f ms' % (d['ms_per_step'], d['value'], d['features_only']['ms_per_step']))
for k,v in d['stages'].items(): print('  %-18s %.3f ms' % (k, v['ms_per_step']))
PY
ncu --set full --clock-control none --import-source on -k regex:k_assign -s 14 -c 1 -o gpurun_out/r02b_assign -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_assign_ncu.log 2>&1; echo "ncu assign rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_update -s 14 -c 1 -o gpurun_out/r02b_update -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_update_ncu.log 2>&1; echo "ncu update rc=$?"
