#!/bin/bash
o=gpurun_out/r02_i3.txt; rm -f $o
run() { echo "== $*" >> $o; env "$@" python bench.py --steps 4 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c110-160 >> $o; }
for i in 1 2; do
run A=1
run OFP_LIB=scripts/variants/libofp_k1_head.so
done
cat $o
