#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/dev_lm_clocks.py 2>&1 | tail -2
ISB_LM_OPERANDS=smem timeout 120 python scripts/dev_lm_clocks.py 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lm_conv_ts -s 2 -c 1 -o gpurun_out/r02_lm_ts python scripts/dev_lm_time.py 2048 > /dev/null 2>&1
ls -la gpurun_out/r02_lm_ts.ncu-rep
