#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dev_umma_rate.py 2>&1 | tail -16
timeout 300 python scripts/dev_banded_lm.py normal 2>&1 | tail -12
timeout 300 python scripts/dev_banded_lm.py short 2>&1 | tail -12
