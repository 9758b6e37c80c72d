#!/bin/bash
# Round-2 batch G: run-pointer representation of the pipelined pass variant; histogram fast path; v1 restored.
mkdir -p gpurun_out
O=gpurun_out/r2g
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=10 > /dev/null
for rep in 1 2; do
echo "== default v1 (rep $rep)"; python tools/quick_bench.py kinds=uniform,range iters=10
for f in variants/*.so; do echo "== V2 $f (rep $rep)"; RSB200_SWEEP_V2=1 RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range iters=10; done
done
echo "== default perm / u64 / u64+perm / i32"; python tools/quick_bench.py kinds=uniform perm=1 iters=10; python tools/quick_bench.py kinds=uniform types=u64 iters=10; python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10
echo "== V2 rp2 perm / u64 / u64+perm"; export RSB200_SWEEP_V2=1 RSB200_LIB=variants/librsb200_U32A_t512_k22_b2_rp2.so; python tools/quick_bench.py kinds=uniform perm=1 iters=10; python tools/quick_bench.py kinds=uniform types=u64 iters=10; python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
