#!/bin/bash
o=gpurun_out/r02_g3.txt; rm -f $o
python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_sizes.py tests/test_gpu_stream_locate.py -x -q 2>&1 | tail -3 >> $o
python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
OFP_K1_WPC=1 python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
python bench.py --steps 3 --warmup 3 --k1-only --no-rel 2>&1 | tail -1 | cut -c1-330 >> $o
cat $o
