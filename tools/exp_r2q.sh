#!/bin/bash
# Round-2 batch Q: next tile staged by per-thread cp.async (LDGSTS) instead of the TMA bulk copy.
mkdir -p gpurun_out
O=gpurun_out/r2q
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=20 > /dev/null
for rep in 1 2; do
echo "== default, TMA bulk copy (rep $rep)"; python tools/quick_bench.py kinds=uniform,range iters=10
echo "== cp.async staging (rep $rep)"; RSB200_LIB=variants/librsb200_U32A_t512_k23_b2_cpasync.so python tools/quick_bench.py kinds=uniform,range iters=10
done
echo "== cp.async perm / u64"; RSB200_LIB=variants/librsb200_U32A_t512_k23_b2_cpasync.so python tools/quick_bench.py kinds=uniform perm=1 iters=10; RSB200_LIB=variants/librsb200_U32A_t512_k23_b2_cpasync.so python tools/quick_bench.py kinds=uniform types=u64 iters=10
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
RSB200_LIB=variants/librsb200_U32A_t512_k23_b2_cpasync.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "keys_only_bit_exact or stable_permutation or size_sweep or unaligned or chunk" 2>&1 | tail -2
