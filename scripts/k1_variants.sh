#!/bin/bash
# Build an A/B variant of onset_detect.cu into scripts/variants/libofp_k1_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/variants
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include \
  $@ -c -o scripts/variants/od_$name.o onset_fingerprinting_b200/csrc/onset_detect.cu
objs=$(ls onset_fingerprinting_b200/csrc/build/*.o | grep -v onset_detect.o)
/usr/local/cuda/bin/nvcc -shared -o scripts/variants/libofp_k1_$name.so $objs scripts/variants/od_$name.o 2>/dev/null
