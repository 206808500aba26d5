#!/bin/bash
python -m pytest tests/test_gpu_detect.py -x -q 2>&1 | tail -3 > gpurun_out/r02_tests_e.txt
for i in 1 2; do
python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-300 >> gpurun_out/r02_tests_e.txt
OFP_LIB=scripts/variants/libofp_k1_r01.so python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-300 >> gpurun_out/r02_tests_e.txt
done
cat gpurun_out/r02_tests_e.txt
