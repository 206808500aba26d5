#!/bin/bash
# ubench (fixed) + ncu source-level profile of k1_detect (grid 1000, 1 s) + traffic at the headline shape
./scripts/ubench/op_rates > gpurun_out/r02_ubench_op_rates.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_detect -c 1 -o gpurun_out/r02_prof_k1_base -f python bench.py --steps 1 --warmup 0 --k1-only --seconds 1 > gpurun_out/r02_ncu_k1_base.log 2>&1
tail -2 gpurun_out/r02_ncu_k1_base.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k1_detect -c 2 --csv --log-file gpurun_out/r02_k1_traffic_base.csv python bench.py --steps 1 --warmup 1 --k1-only > gpurun_out/r02_ncu_k1_traffic.log 2>&1
tail -3 gpurun_out/r02_k1_traffic_base.csv
grep -v "^FADD\|^FFMA" gpurun_out/r02_ubench_op_rates.txt | grep "SMSP=2"
