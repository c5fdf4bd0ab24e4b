#!/bin/bash
# dev: region growing, median, morphology on the GPU + config 4 / config 3 bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_region_growing.py tests/test_gpu_descriptor_drivers.py tests/test_gpu_round2.py -q > gpurun_out/r02_rg_tests.log 2>&1; echo "rg tests rc=$?"; tail -40 gpurun_out/r02_rg_tests.log
timeout 600 python bench.py --workload config4 --steps 3 --warmup 1 > gpurun_out/r02_config4.json 2> gpurun_out/r02_config4.err; echo "config4 rc=$?"; tail -c 1500 gpurun_out/r02_config4.json; tail -5 gpurun_out/r02_config4.err
timeout 600 python bench.py --workload config3 --steps 5 --warmup 2 > gpurun_out/r02_config3_n1.json 2> gpurun_out/r02_config3_n1.err; echo "config3 rc=$?"; tail -c 1800 gpurun_out/r02_config3_n1.json; tail -5 gpurun_out/r02_config3_n1.err
