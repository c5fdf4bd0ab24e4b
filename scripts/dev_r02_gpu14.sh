#!/bin/bash
mkdir -p gpurun_out
echo "== LM, merged-N TS, warp-wide TMA issue"
timeout 300 python scripts/dev_lm_time.py 2048 | tail -3
timeout 900 python -m pytest tests/test_gpu_texture.py tests/test_gpu_round2.py -q -x 2>&1 | tail -3
