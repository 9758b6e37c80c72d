#!/bin/bash
# Round-2 batch O: 32-bit run starts with the rows kept in order (RS_X_OFF32) on the new default; tile shapes of the
# merge tile kernel on local runs (2^30 keys).
mkdir -p gpurun_out
O=gpurun_out/r2o
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=20 > /dev/null
for rep in 1 2; do
echo "== default (rep $rep)"; python tools/quick_bench.py kinds=uniform,range iters=10
for f in variants/*off32*.so; do echo "== $f (rep $rep)"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range iters=10; done
done
echo "== off32 perm / u64"; RSB200_LIB=variants/librsb200_U32A_t512_k23_b2_off32.so python tools/quick_bench.py kinds=uniform perm=1 iters=10; RSB200_LIB=variants/librsb200_U32A_t512_k23_b2_off32.so python tools/quick_bench.py kinds=uniform types=u64 iters=10
for rep in 1 2; do
echo "== merge default 256x16x4 (rep $rep)"; python tools/merge_bench.py log2n=30 iters=5; python tools/merge_bench.py log2n=30 iters=5 jobs=4
for f in variants/*_m*.so; do echo "== $f (rep $rep)"; RSB200_LIB=$f python tools/merge_bench.py log2n=30 iters=5; RSB200_LIB=$f python tools/merge_bench.py log2n=30 iters=5 jobs=4; done
done
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
