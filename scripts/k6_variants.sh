#!/bin/bash
# Build an A/B variant of cnn_infer.cu into scripts/variants/libofp_k6_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/variants
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include \
  $@ -Xptxas=-v -c -o scripts/variants/cnn_$name.o onset_fingerprinting_b200/csrc/cnn_infer.cu 2>&1 | grep -A3 "k6_cccnn_ctaILi3ELi2" | grep -E "Used|spill"
objs=$(ls onset_fingerprinting_b200/csrc/build/*.o | grep -v cnn_infer.o)
/usr/local/cuda/bin/nvcc -shared -o scripts/variants/libofp_k6_$name.so $objs scripts/variants/cnn_$name.o 2>/dev/null
