#!/bin/bash
o=gpurun_out/r02_ab.txt; rm -f $o
python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -2 >> $o
run() { echo "== $*" >> $o; env "$@" python bench.py --steps 4 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c110-160 >> $o; }
for i in 1 2 3; do
run A=1
run OFP_LIB=scripts/variants/libofp_k1_head.so
done
cat $o
