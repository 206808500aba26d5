#!/bin/bash
python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_sizes.py -x -q 2>&1 | tail -5 > gpurun_out/r02_tests_c.txt
run() { echo "== $*" >> gpurun_out/r02_k1_split.txt; env "$@" python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-400 >> gpurun_out/r02_k1_split.txt; }
rm -f gpurun_out/r02_k1_split.txt
run OFP_K1_SPLIT=1
run OFP_K1_SPLIT=1 OFP_K1_ROLESWAP=1
run OFP_K1_SPLIT=0
OFP_K1_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:k1_detect2 -c 1 -o gpurun_out/r02_prof_k1_split -f python bench.py --steps 1 --warmup 0 --k1-only --seconds 1 > gpurun_out/r02_ncu_split.log 2>&1
cat gpurun_out/r02_tests_c.txt gpurun_out/r02_k1_split.txt
