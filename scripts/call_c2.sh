#!/bin/bash
o=gpurun_out/r02_c2.txt; rm -f $o
for v in "" kqd ku4; do
  if [ -z "$v" ]; then lib=""; else lib="OFP_LIB=scripts/variants/libofp_k1_$v.so"; fi
  echo "== ${v:-default}" >> $o
  env $lib python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
done
cat $o
ncu --set full --clock-control none --import-source on -k regex:k4_fix -c 1 -o gpurun_out/r02_prof_k4_a -f python bench.py --workload hits16 --hits 40000 --steps 1 --warmup 0 --skip-cpu --skip-e2e > gpurun_out/r02_ncu_k4_a.log 2>&1
tail -2 gpurun_out/r02_ncu_k4_a.log
