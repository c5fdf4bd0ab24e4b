#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dev_lm_time.py 2048 | tail -3
timeout 900 python -m pytest tests/test_gpu_texture.py tests/test_gpu_round2.py tests/test_gpu_tiled.py tests/test_reference_vectors.py -q -x 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_lm_launches.csv python scripts/dev_lm_time.py 2048 > /dev/null 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open("gpurun_out/r02b_lm_launches.csv") if not l.startswith("==")]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name")!="gpu__time_duration.sum": continue
    v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
    v = v/1e3 if u=="ns" else (v if u=="us" else v*1e3)
    k=row["Kernel Name"].replace("<unnamed>::","").split("(")[0][:50]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]: print("%-52s n=%4d  %8.1f us each  total %8.2f ms" % (k,c,v/c,v/1e3))
PY
