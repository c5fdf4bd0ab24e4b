#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiled.py tests/test_gpu_volume.py -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f (%.3f ms) e2e %.1f batch %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_batch']['value'])); print({k: round(v['ms_per_step'],4) for k,v in d['stages'].items()})"
timeout 300 python scripts/dev_lm_time.py 2048 | tail -10
