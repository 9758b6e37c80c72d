#!/bin/bash
# Round-2 experiment batch A (run on the GPU box through gpurun): tests with the second-generation sweep kernel,
# A/B against the first generation, tile-shape variants, one ncu capture.
mkdir -p gpurun_out
O=gpurun_out/r2a
timeout 600 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
{
echo "== default (sweep v2, 512x22x2)"; python tools/quick_bench.py kinds=uniform,random,range,inverted,zeros
echo "== v1 kernel"; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform,range
for f in variants/*.so; do echo "== $f"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range; done
echo "== perm / u64"; python tools/quick_bench.py kinds=uniform perm=1; python tools/quick_bench.py kinds=uniform types=u64; python tools/quick_bench.py kinds=uniform types=u64 perm=1
echo "== v1 perm / u64"; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform perm=1; RSB200_SWEEP_V1=1 python tools/quick_bench.py kinds=uniform types=u64
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
ncu --set full --import-source on --clock-control none -k regex:"rs_sweep" -s 4 -c 1 -o ${O}_sweep python tools/quick_bench.py kinds=uniform iters=1 > ${O}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -5
