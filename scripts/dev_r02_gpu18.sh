#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29402 tests/run_tiled_ranks.py > gpurun_out/r02b_tiled_ranks_n2.log 2>&1
echo "tiled parity N=2 rc=$?"; grep -E "TILED-RANKS-OK|Error|assert" gpurun_out/r02b_tiled_ranks_n2.log | head -5; tail -3 gpurun_out/r02b_tiled_ranks_n2.log
