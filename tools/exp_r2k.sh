#!/bin/bash
# Round-2 batch K: fourth-generation pass kernel (RSB200_PASS=4, rs_pass4.cuh: un-serialised shared-memory accesses) and
# the third generation with the same un-serialisation, against the default, interleaved twice.
mkdir -p gpurun_out
O=gpurun_out/r2k
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=20 > /dev/null
for rep in 1 2; do
echo "== v1 default (rep $rep)"; python tools/quick_bench.py kinds=uniform,range iters=10
echo "== v4 RSB200_PASS=4 sb4 (rep $rep)"; RSB200_PASS=4 python tools/quick_bench.py kinds=uniform,range,inverted iters=10
echo "== v3 RSB200_PASS=3 (rep $rep)"; RSB200_PASS=3 python tools/quick_bench.py kinds=uniform,range iters=10
for f in variants/*.so; do echo "== v4 $f (rep $rep)"; RSB200_PASS=4 RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range iters=10; done
done
echo "== v4 zeros without pass skipping does not exist in quick_bench; perm / u64 / u64+perm:"
echo "== v1 perm / u64 / u64+perm"; python tools/quick_bench.py kinds=uniform,range perm=1 iters=10; python tools/quick_bench.py kinds=uniform,range types=u64 iters=10; python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10
echo "== v4 perm / u64 / u64+perm"; export RSB200_PASS=4; python tools/quick_bench.py kinds=uniform,range perm=1 iters=10; python tools/quick_bench.py kinds=uniform,range types=u64 iters=10; python tools/quick_bench.py kinds=uniform types=u64 perm=1 iters=10
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-220
export RSB200_PASS=4
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "keys_only_bit_exact or stable_permutation or size_sweep or golden or bit_ranges or unaligned or float_keys_native or 16_bit or chunk or exact_2_28_u32_key_index or full_size_2_28_u64_key_value or constant_digit" > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest.log
