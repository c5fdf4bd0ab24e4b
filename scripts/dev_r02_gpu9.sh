#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiled.py -q -x 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.1f (%.3f ms) e2e %.1f batch %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_batch']['value'])); print({k: round(v['ms_per_step'],4) for k,v in d['stages'].items()})"
python scripts/dev_gmm_big.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_gmm_big_launches.csv python scripts/dev_gmm_big.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open("gpurun_out/r02_gmm_big_launches.csv") if not l.startswith("==")]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name")!="gpu__time_duration.sum": continue
    v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
    v = v/1e3 if u=="ns" else (v if u=="us" else v*1e3)
    k=row["Kernel Name"].replace("<unnamed>::","").split("(")[0][:50]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-52s n=%4d  %8.1f us each  total %8.2f ms" % (k,c,v/c,v/1e3))
PY
timeout 600 python -m pytest tests/test_gpu_texture.py tests/test_reference_vectors.py tests/test_gpu_round2.py -q -x 2>&1 | tail -2
python scripts/dev_lm_time.py 2048 | tail -9
