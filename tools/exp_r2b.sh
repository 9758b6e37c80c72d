#!/bin/bash
# Round-2 experiment batch B: look-back timing x table packing x window, new parity tests, ncu of the default.
mkdir -p gpurun_out
O=gpurun_out/r2b
{
echo "== default lib"; python tools/quick_bench.py kinds=uniform,range,zeros
echo "== v1 kernel"; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform
for f in variants/*.so; do echo "== $f"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range; done
echo "== perm / u64"; python tools/quick_bench.py kinds=uniform perm=1; python tools/quick_bench.py kinds=uniform types=u64; python tools/quick_bench.py kinds=uniform types=u64 perm=1
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest.log
ncu --set full --import-source on --clock-control none -k regex:"rs_sweep" -s 4 -c 1 -o ${O}_sweep python tools/quick_bench.py kinds=uniform iters=1 > ${O}_ncu.log 2>&1; echo "ncu rc=$?"
