#!/bin/bash
o=gpurun_out/r02_c4.txt; rm -f $o
for v in launder nokmagic; do
  echo "== $v" >> $o
  OFP_LIB=scripts/variants/libofp_k1_$v.so python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
done
cat $o
