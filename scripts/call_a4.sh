#!/bin/bash
out=gpurun_out/r02_a4.txt; rm -f $out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_stream_locate.py -x -q -k "graph_equals" 2>&1 | grep -v "^$" | head -60 > gpurun_out/r02_sanitizer.txt
for v in "" noicvt nokmagic nofolmax nomnvote; do
  if [ -z "$v" ]; then lib=""; else lib="OFP_LIB=scripts/variants/libofp_k1_$v.so"; fi
  echo "== ${v:-default}" >> $out
  env $lib python bench.py --steps 3 --warmup 3 --k1-only 2>&1 | tail -1 | cut -c1-330 >> $out
done
cat $out; head -40 gpurun_out/r02_sanitizer.txt
