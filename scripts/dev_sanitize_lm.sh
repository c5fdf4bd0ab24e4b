#!/bin/bash
timeout 110 compute-sanitizer --tool memcheck --print-limit 20 python scripts/dev_sanitize_lm.py 2>&1 | grep -v "^$" | tail -14
