#!/bin/bash
o=gpurun_out/r02_b3.txt; rm -f $o
python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_sizes.py tests/test_gpu_stream_locate.py -x -q 2>&1 | tail -3 >> $o
for v in "" noskipfol nomxspec; do
  if [ -z "$v" ]; then lib=""; else lib="OFP_LIB=scripts/variants/libofp_k1_$v.so"; fi
  echo "== ${v:-default}" >> $o
  env $lib python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
done
cat $o
