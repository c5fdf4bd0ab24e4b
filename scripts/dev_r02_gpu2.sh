#!/bin/bash
# dev: launch list of the texture pipeline and one full ncu capture of the tcgen05 contraction
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_lm_launches.csv python scripts/dev_lm_time.py 2048 > gpurun_out/r02_lm_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_lm_conv_tc -s 1 -c 1 -o gpurun_out/r02_lm_conv_tc -f python scripts/dev_lm_time.py 2048 > gpurun_out/r02_lm_ncu.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -5
