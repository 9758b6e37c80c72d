#!/bin/bash
# Round-2 batch M: single-change variants of the default pass kernel (variants/*.so, see the RS_X_* switches in
# rs_kernels.cuh); default first and last.
mkdir -p gpurun_out
O=gpurun_out/r2m
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=20 > /dev/null
echo "== default (first)"; python tools/quick_bench.py kinds=uniform,range iters=10
for f in variants/*.so; do echo "== $f"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range iters=10; done
echo "== default (last)"; python tools/quick_bench.py kinds=uniform,range iters=10
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
