#!/bin/bash
run() { python bench.py --workload hits16 --hits 100000 --steps 3 --warmup 2 --skip-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), 'ms', round(d['value']/1e6,3), 'Mhits/s fix_ok', d['fix_ok'], d['located'])"; }
run default
for f in scripts/variants/libofp_k4_*.so; do OFP_LIB=$f python -m pytest tests/test_gpu_lag_locate.py -x -q 2>&1 | tail -1; OFP_LIB=$f run $(basename $f); done
