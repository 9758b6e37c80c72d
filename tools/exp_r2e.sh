#!/bin/bash
# Round-2 experiment batch E: warp-cooperative look-back.  Every configuration is run twice, interleaved (the first
# process on a fresh box was seen to run ~13 % slow), uniform + range.
mkdir -p gpurun_out
O=gpurun_out/r2e
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=10 > /dev/null
for rep in 1 2; do
echo "== v1 kernel (rep $rep)"; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform,range iters=10
echo "== default lib (rep $rep)"; python tools/quick_bench.py kinds=uniform,range iters=10
for f in variants/*.so; do echo "== $f (rep $rep)"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range iters=10; done
done
echo "== all datasets default"; python tools/quick_bench.py iters=10
echo "== perm / u64"; python tools/quick_bench.py kinds=uniform perm=1 iters=10; python tools/quick_bench.py kinds=uniform types=u64 iters=10; python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10
echo "== v1 perm / u64"; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform perm=1 iters=10; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform types=u64 iters=10;  RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest.log
ncu --set full --import-source on --clock-control none -k regex:"rs_sweep" -s 4 -c 1 -o ${O}_sweep python tools/quick_bench.py kinds=uniform iters=1 > ${O}_ncu.log 2>&1; echo "ncu rc=$?"
