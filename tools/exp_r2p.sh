#!/bin/bash
# Round-2 batch P: split-half histogram for 8-byte keys (lanes 0-15 count the low word, lanes 16-31 the high word).
mkdir -p gpurun_out
O=gpurun_out/r2p
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=20 > /dev/null
echo "== u64 / i64 / u64+perm (split-half histogram)"; python tools/quick_bench.py kinds=uniform,range,zeros types=u64 iters=10; python tools/quick_bench.py kinds=uniform types=i64 iters=10; python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10
echo "== u32 (unchanged)"; python tools/quick_bench.py kinds=uniform iters=10
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-220
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
