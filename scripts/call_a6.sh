#!/bin/bash
o=gpurun_out/r02_a6.txt; rm -f $o
for i in 1 2 3; do python -m pytest tests/test_gpu_stream_locate.py -x -q 2>&1 | tail -2 >> $o; done
echo "== head lib" >> $o
for i in 1 2 3; do OFP_LIB=scripts/variants/libofp_k1_head.so python -m pytest tests/test_gpu_stream_locate.py -x -q 2>&1 | tail -2 >> $o; done
echo "== alone" >> $o
for i in 1 2; do python -m pytest tests/test_gpu_stream_locate.py -x -q -k graph_equals 2>&1 | tail -2 >> $o; done
cat $o
