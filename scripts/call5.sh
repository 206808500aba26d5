#!/bin/bash
python -m pytest tests/test_gpu_stream_locate.py tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_sizes.py tests/test_gpu_lag_locate.py tests/test_gpu_tools.py -x -q 2>&1 | tail -25 > gpurun_out/r02_tests_d.txt
python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-300 > gpurun_out/r02_k1_trim.txt
cat gpurun_out/r02_tests_d.txt gpurun_out/r02_k1_trim.txt
