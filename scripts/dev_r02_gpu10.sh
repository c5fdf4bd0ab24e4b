#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py -q 2>&1 | tail -4
echo "== LM, A operand in tensor memory (default)"
timeout 300 python scripts/dev_lm_time.py 2048 | tail -12
echo "== LM, both operands in shared memory"
ISB_LM_OPERANDS=smem timeout 300 python scripts/dev_lm_time.py 2048 | tail -12
timeout 900 python -m pytest tests/test_gpu_texture.py tests/test_gpu_round2.py tests/test_gpu_tiled.py tests/test_reference_vectors.py -q -x 2>&1 | tail -6
