#!/bin/bash
o=gpurun_out/r02_a7.txt; rm -f $o
CUDA_LAUNCH_BLOCKING=1 python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_sizes.py tests/test_gpu_stream_locate.py -x -q 2>&1 | tail -60 >> $o
echo "== nonblocking full" >> $o
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $o
cat $o
