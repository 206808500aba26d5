#!/bin/bash
o=gpurun_out/r02_b7.txt; rm -f $o
run() { echo "== $*" >> $o; env "$@" python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o; }
run OFP_K1_PAD=0
run OFP_K1_PAD=1 OFP_K1_TILE=32
run OFP_K1_PAD=0 OFP_K1_TILE=32
run OFP_K1_PAD=1 OFP_K1_TILE=24
run OFP_K1_PAD=1 OFP_K1_TILE=40
OFP_K1_PAD=1 OFP_K1_TILE=32 python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3 >> $o
cat $o
