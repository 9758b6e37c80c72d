#!/bin/bash
# Developer A/B builds of librsb200.so with different u32 keys-only tile shapes -> gpurun_variants/*.so
# usage: tools/build_variants.sh "256 32 2" "256 16 4" ...   (THREADS KPT MIN_BLOCKS)
set -e
cd "$(dirname "$0")/../radix-sort_b200/csrc"
mkdir -p ../../variants
for cfg in "$@"; do
  set -- $cfg
  out=../../variants/librsb200_t$1_k$2_b$3.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 177 \
       -DRS_U32_THREADS=$1 -DRS_U32_KPT=$2 -DRS_U32_MINB=$3 $EXTRA -Xptxas -v -shared -o $out rs_api.cu 2> /tmp/variant_ptxas.log
  echo "$out: $(grep -A2 'onesweep_kernelIjLb0E' /tmp/variant_ptxas.log | grep -E 'registers|spill' | sed 's/ptxas info    : //' | tr '\n' ' ')"
done
