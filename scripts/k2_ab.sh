#!/bin/bash
# A/B of K2 variants on the GPU box: default library first, then every scripts/variants/libofp_*.so
run() { python bench.py --workload spectral --recordings 2000 --steps 3 --warmup 3 --skip-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), d['roofline']['note'])"; }
python -m pytest tests/test_gpu_spectral.py -x -q 2>&1 | tail -1
run default
for f in scripts/variants/libofp_*.so; do OFP_LIB=$f run $(basename $f); done
