#!/bin/bash
mkdir -p gpurun_out
echo "== LM, merged-N, A operand in tensor memory (default)"
timeout 300 python scripts/dev_lm_time.py 2048 | tail -3
echo "== LM, both operands in shared memory"
ISB_LM_OPERANDS=smem timeout 300 python scripts/dev_lm_time.py 2048 | tail -3
timeout 900 python -m pytest tests/test_gpu_texture.py tests/test_gpu_round2.py tests/test_gpu_tiled.py tests/test_reference_vectors.py -q -x 2>&1 | tail -25
