#!/bin/bash
# dev: first GPU pass of round 2 -- tcgen05 self-test, texture parity, LM timing, the whole gpu suite, the bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_umma.py -q > gpurun_out/r02_umma.log 2>&1; echo "umma rc=$?"; tail -15 gpurun_out/r02_umma.log
timeout 600 python -m pytest tests/test_gpu_texture.py tests/test_reference_vectors.py -q > gpurun_out/r02_texture.log 2>&1; echo "texture rc=$?"; tail -25 gpurun_out/r02_texture.log
timeout 300 python scripts/dev_lm_time.py > gpurun_out/r02_lm_time.log 2>&1; echo "lm_time rc=$?"; tail -20 gpurun_out/r02_lm_time.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -30 gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r02_bench1.json; tail -5 gpurun_out/r02_bench1.err
