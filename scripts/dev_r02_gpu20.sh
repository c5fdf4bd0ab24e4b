#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_region_growing.py -q -x 2>&1 | tail -3
