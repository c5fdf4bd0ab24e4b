#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dev_lm_time.py 2048 | tail -3
timeout 900 python -m pytest tests/test_gpu_texture.py tests/test_gpu_round2.py tests/test_gpu_tiled.py tests/test_reference_vectors.py -q -x 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none -k regex:k_lm_vblur -s 4 -c 1 -o gpurun_out/r02_lm_vblur python scripts/dev_lm_time.py 2048 > /dev/null 2>&1
ncu -i gpurun_out/r02_lm_vblur.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for a,b,c in zip(h,u,v):
    if any(k in a for k in ('gpu__time_duration.sum','sm__inst_executed_pipe_fp64','pipe_fp64_cycles_active','lts__throughput.avg.pct','l1tex__m_xbar2l1tex_read_bytes.sum ','l1tex__m_xbar2l1tex_read_bytes.sum.per_second','dram__bytes_read.sum ','sm__throughput.avg.pct','smsp__issue_active.avg.per_cycle_active','sm__warps_active.avg.pct','l1tex__t_sector_hit_rate')): print(a,b,c)
"
