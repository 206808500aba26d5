#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:k1_detect -c 1 -o gpurun_out/r02_prof_k1_e -f python bench.py --steps 1 --warmup 0 --k1-only --seconds 1 > gpurun_out/r02_ncu_k1_e.log 2>&1
tail -1 gpurun_out/r02_ncu_k1_e.log
