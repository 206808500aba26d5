#!/bin/bash
o=gpurun_out/r02_c3.txt; rm -f $o
for v in "" icin icout icboth; do
  if [ -z "$v" ]; then lib=""; else lib="OFP_LIB=scripts/variants/libofp_k1_$v.so"; fi
  echo "== ${v:-default}" >> $o
  env $lib python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $o
done
OFP_LIB=scripts/variants/libofp_k1_icboth.so python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_sizes.py -x -q 2>&1 | tail -2 >> $o
python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py tests/test_gpu_sizes.py -x -q 2>&1 | tail -2 >> $o
cat $o
