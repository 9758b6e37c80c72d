#!/bin/bash
# Round-2 batch L: two look-back windows per round trip (RS_LB2: both halves of the CTA read one) + cached chunk base,
# on the default (first-generation) kernel, window sizes 4 / 8 / 12 / 16; interleaved twice.
mkdir -p gpurun_out
O=gpurun_out/r2l
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=20 > /dev/null
for rep in 1 2; do
echo "== default (rep $rep)"; python tools/quick_bench.py kinds=uniform,range iters=10
for f in variants/*.so; do echo "== $f (rep $rep)"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range iters=10; done
done
export RSB200_LIB=variants/librsb200_U32A_t512_k22_b2_lb2.so
echo "== lb2 perm / u64 / u64+perm"; python tools/quick_bench.py kinds=uniform,range perm=1 iters=10; python tools/quick_bench.py kinds=uniform,range types=u64 iters=10; python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-220
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "keys_only_bit_exact or stable_permutation or size_sweep or golden or bit_ranges or unaligned or float_keys_native or 16_bit or chunk or exact_2_28_u32_key_index or full_size_2_28_u64_key_value or constant_digit or ranking_modes" > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest.log
