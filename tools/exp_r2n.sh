#!/bin/bash
# Round-2 batch N: new default (u32 tiles of 512 x 23 keys + cached chunk base) against 512 x 22 + cached base; then the
# whole GPU suite, the contract bench (both arms) and the profiles of the new default.
mkdir -p gpurun_out
O=gpurun_out/r2n
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=20 > /dev/null
for rep in 1 2; do
echo "== default 512x23 + cbase (rep $rep)"; python tools/quick_bench.py iters=10
echo "== 512x22 + cbase (rep $rep)"; RSB200_LIB=variants/librsb200_U32A_t512_k22_b2.so python tools/quick_bench.py iters=10
done
echo "== default perm / u64 / u64+perm / i32"; python tools/quick_bench.py kinds=uniform,range perm=1 iters=10; python tools/quick_bench.py kinds=uniform,range types=u64 iters=10; python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10; python tools/quick_bench.py kinds=uniform,inverted types=i32 iters=10
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 ${O}_pytest.log
python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; cut -c1-400 ${O}_bench.json
bash tools/gpu_profile.sh r2n
