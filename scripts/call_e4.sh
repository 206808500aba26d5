#!/bin/bash
o=gpurun_out/r02_e4.txt; rm -f $o
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $o
python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
python bench.py --workload realtime --steps 3 --warmup 3 > gpurun_out/r02_bench_realtime.json 2>> $o; cut -c1-900 gpurun_out/r02_bench_realtime.json >> $o
cat $o
