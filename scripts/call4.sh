#!/bin/bash
OFP_K1_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:k1_detect2 -c 1 -o gpurun_out/r02_prof_k1_split -f python bench.py --steps 1 --warmup 0 --k1-only --seconds 1 > gpurun_out/r02_ncu_split.log 2>&1
tail -3 gpurun_out/r02_ncu_split.log
