#!/bin/bash
o=gpurun_out/r02_e5.txt; rm -f $o
python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_stream_locate.py -x -q 2>&1 | tail -2 >> $o
python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
python bench.py --workload realtime --steps 3 --warmup 3 > gpurun_out/r02_bench_realtime.json 2>> $o; cut -c1-700 gpurun_out/r02_bench_realtime.json >> $o
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_rt.csv python bench.py --workload realtime --blocks 40 > gpurun_out/r02_ncu_rt.log 2>&1
python - <<'PY' >> $o
import csv,collections
rows=list(csv.DictReader(l for l in open('gpurun_out/r02_launches_rt.csv') if l.startswith('"')))
agg=collections.defaultdict(list)
for r in rows: agg[r['Kernel Name'][:50]].append(float(r['Metric Value'])/1e3)
for k,v in agg.items(): print(f"{len(v):4d}x median {sorted(v)[len(v)//2]:8.1f} us  max {max(v):8.1f}  {k}")
PY
cat $o
