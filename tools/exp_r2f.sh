#!/bin/bash
# Round-2 validation batch (1 GPU): full GPU test suite, the contract bench (both arms), profiles, sanitizer.
mkdir -p gpurun_out
O=gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 ${O}_pytest.log
python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; cut -c1-600 ${O}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; echo "bench ref rc=$?"; cut -c1-400 ${O}_bench_ref.json
{
for rep in 1 2; do
echo "== default (v1 + 32-bit offsets) rep $rep"; python tools/quick_bench.py kinds=uniform,range iters=10
echo "== RSB200_SWEEP_V2=1 rep $rep"; RSB200_SWEEP_V2=1 python tools/quick_bench.py kinds=uniform,range iters=10
done
echo "== default all datasets"; python tools/quick_bench.py iters=10
echo "== default perm / u64 / u64+perm"; python tools/quick_bench.py kinds=uniform perm=1 iters=10; python tools/quick_bench.py kinds=uniform types=u64 iters=10; python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
bash tools/gpu_profile.sh r2f
bash tools/sanitize.sh r2f
