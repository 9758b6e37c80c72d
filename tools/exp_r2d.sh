#!/bin/bash
# Round-2 experiment batch D: the software-pipelined sweep kernel against the first generation, window / tile variants.
mkdir -p gpurun_out
O=gpurun_out/r2d
{
echo "== default lib (pipelined)"; python tools/quick_bench.py kinds=uniform,random,range,inverted,zeros
echo "== v1 kernel"; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform,range
for f in variants/*.so; do echo "== $f"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range; done
echo "== perm / u64 / f32 / u16"; python tools/quick_bench.py kinds=uniform perm=1; python tools/quick_bench.py kinds=uniform types=u64; python tools/quick_bench.py kinds=uniform types=u64 perm=1
echo "== v1 perm / u64"; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform perm=1; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform types=u64;  RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform types=u64 perm=1
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest.log
ncu --set full --import-source on --clock-control none -k regex:"rs_sweep" -s 4 -c 1 -o ${O}_sweep python tools/quick_bench.py kinds=uniform iters=1 > ${O}_ncu.log 2>&1; echo "ncu rc=$?"
