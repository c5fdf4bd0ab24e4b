#!/bin/bash
# dev: source-level ncu captures of the two SLIC sweep kernels on the bench image + the large-image LM parity test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q > gpurun_out/r02_round2_tests.log 2>&1; echo "round2 tests rc=$?"; tail -15 gpurun_out/r02_round2_tests.log
ncu --set full --clock-control none --import-source on -k regex:k_assign -s 14 -c 1 -o gpurun_out/r02_assign -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_assign_ncu.log 2>&1; echo "ncu assign rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_update -s 14 -c 1 -o gpurun_out/r02_update -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_update_ncu.log 2>&1; echo "ncu update rc=$?"
ls -la gpurun_out/*.ncu-rep
