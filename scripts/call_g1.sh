#!/bin/bash
o=gpurun_out/r02_g1.txt; rm -f $o
python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -2 >> $o
python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
cat $o
