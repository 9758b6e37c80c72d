#!/bin/bash
# Developer A/B builds of librsb200.so with different tile shapes of the atomic ranking mode -> variants/*.so
# usage: tools/build_variants.sh "U32A 512 20 2" "U32VA 512 14 2" "U64A 512 10 2" ...   (CFG THREADS KPT MIN_BLOCKS)
# EXTRA="-DRS_LB_WINDOW=16" adds defines; TAG=name is appended to the file names.  Builds run in parallel.
set -e
cd "$(dirname "$0")/../radix-sort_b200/csrc"
mkdir -p ../../variants
pids=()
for cfg in "$@"; do
  set -- $cfg
  out=../../variants/librsb200_$1_t$2_k$3_b$4${TAG:+_$TAG}.so
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 177 \
       -DRS_$1_THREADS=$2 -DRS_$1_KPT=$3 -DRS_$1_MINB=$4 $EXTRA -Xptxas -v -shared -o $out rs_api.cu 2> $out.ptxas.log
    echo "$out: $(grep -A2 "rs_sweep_kernel.*ELi$2ELi$3ELi$4E" $out.ptxas.log | grep -E 'registers|spill' | sed 's/ptxas info    : //' | tr '\n' ' ' | cut -c1-300)" ) &
  pids+=($!)
  if [ ${#pids[@]} -ge 6 ]; then wait ${pids[0]}; pids=("${pids[@]:1}"); fi
done
wait
# Merge tile shapes: EXTRA="-DRS_MERGE_THREADS=128 -DRS_MERGE_ITEMS=32 -DRS_MERGE_MINB32=6" tools/build_variants.sh "U32A 512 22 2"
# (RS_MERGE_MINB32 / RS_MERGE_MINB32_DENSE / RS_MERGE_MINB64 = resident CTAs per SM the tile kernel is compiled for);
# run with RSB200_LIB=variants/<file>.so python tools/merge_bench.py
