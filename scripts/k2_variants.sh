#!/bin/bash
# Build A/B variants of the K2 warp kernel into scripts/variants/ (git-ignored .so, travel with gpurun).
# usage: scripts/k2_variants.sh name "-DOFP_K2W_WARPS=6 -DOFP_K2W_FRAMES=8"
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/variants
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include \
  $@ -Xptxas=-v -c -o scripts/variants/sf_$name.o onset_fingerprinting_b200/csrc/spectral_flux.cu 2>&1 | grep -A2 "k2_flux_warpILi0" | grep -E "Used|spill"
objs=$(ls onset_fingerprinting_b200/csrc/build/*.o | grep -v spectral_flux.o)
/usr/local/cuda/bin/nvcc -shared -o scripts/variants/libofp_$name.so $objs scripts/variants/sf_$name.o
