#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gmm.py tests/test_gpu_round2.py -q -x 2>&1 | tail -3
python scripts/dev_gmm_big.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_gmm_big_launches.csv python scripts/dev_gmm_big.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open("gpurun_out/r02b_gmm_big_launches.csv") if not l.startswith("==")]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name")!="gpu__time_duration.sum": continue
    v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
    v = v/1e3 if u=="ns" else (v if u=="us" else v*1e3)
    k=row["Kernel Name"].replace("<unnamed>::","").split("(")[0][:50]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-52s n=%4d  %8.1f us each  total %8.2f ms" % (k,c,v/c,v/1e3))
PY
python bench.py --workload config3 --steps 6 --warmup 2 > gpurun_out/r02_config3_n1.json 2> gpurun_out/r02_config3_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_config3_n1.json').read().strip().splitlines()[-1]); print('config3 %.2f ms' % d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['stages'].items() if v['ms_per_step']>1})"
