#!/bin/bash
python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_sizes.py tests/test_gpu_stream_locate.py -x -q 2>&1 | tail -15 > gpurun_out/r02_tests_a3.txt
python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-400 >> gpurun_out/r02_tests_a3.txt
OFP_LIB=scripts/variants/libofp_k1_nosparse.so python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-400 >> gpurun_out/r02_tests_a3.txt
cat gpurun_out/r02_tests_a3.txt
