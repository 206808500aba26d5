#!/bin/bash
# Build an A/B variant of lag_fix.cu into scripts/variants/libofp_k4_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/variants
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include \
  $@ -c -o scripts/variants/lf_$name.o onset_fingerprinting_b200/csrc/lag_fix.cu
objs=$(ls onset_fingerprinting_b200/csrc/build/*.o | grep -v lag_fix.o)
/usr/local/cuda/bin/nvcc -shared -o scripts/variants/libofp_k4_$name.so $objs scripts/variants/lf_$name.o 2>/dev/null
