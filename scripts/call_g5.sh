#!/bin/bash
o=gpurun_out/r02_g5.txt; rm -f $o
run() { echo "== $*" >> $o; env "$@" python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c60-200 >> $o; }
run OFP_K1_STAGGER_NS=0
run OFP_K1_STAGGER_NS=150
run OFP_K1_STAGGER_NS=300
run OFP_K1_STAGGER_NS=600
run OFP_K1_STAGGER_NS=300 OFP_K1_STAGGER_FROM=500
cat $o
