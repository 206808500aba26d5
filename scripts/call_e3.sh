#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_rt.csv python bench.py --workload realtime --blocks 60 > gpurun_out/r02_ncu_rt.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.DictReader(l for l in open('gpurun_out/r02_launches_rt.csv') if l.startswith('"')))
agg=collections.defaultdict(list)
for r in rows: agg[r['Kernel Name'][:50]].append(float(r['Metric Value'])/1e3)
for k,v in agg.items(): print(f"{len(v):4d}x median {sorted(v)[len(v)//2]:8.1f} us  max {max(v):8.1f}  {k}")
PY
