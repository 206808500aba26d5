#!/bin/bash
o=gpurun_out/r02_g2.txt; rm -f $o
python -m pytest tests/test_gpu_lag_locate.py tests/test_gpu_sizes.py tests/test_gpu_pipeline.py tests/test_gpu_tools.py -x -q 2>&1 | tail -2 >> $o
python bench.py --workload hits16 --hits 200000 --steps 3 --warmup 2 --skip-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), 'ms', round(d['value']/1e6,3), 'Mhits/s fix_ok', d['fix_ok'], d['located'], d['parity_sample']['ok'], d['cc_screening'])" >> $o
cat $o
