#!/bin/bash
CUDA_LAUNCH_BLOCKING=1 python -m pytest tests/test_gpu_stream_locate.py -x -q 2>&1 | tail -40 > gpurun_out/r02_a5.txt
CUDA_LAUNCH_BLOCKING=1 python -m pytest tests/test_gpu_stream_locate.py -x -q -k "graph_equals" 2>&1 | tail -40 >> gpurun_out/r02_a5.txt
cat gpurun_out/r02_a5.txt
