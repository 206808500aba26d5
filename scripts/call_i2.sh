#!/bin/bash
o=gpurun_out/r02_i2.txt; rm -f $o
python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_stream_locate.py -x -q 2>&1 | tail -2 >> $o
python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --skip-hits16 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step ms',d['ms_per_step'],'k1',d['roofline']['kernel_ms'],'value',d['value'])" >> $o
cat $o
