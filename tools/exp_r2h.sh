#!/bin/bash
# Round-2 batch H: tile shapes of the key+index and 64-bit instantiations (first-generation kernel = default); own bounds checks.
mkdir -p gpurun_out
O=gpurun_out/r2h
{
echo "== warm-up process (ignored)"; python tools/quick_bench.py kinds=uniform iters=10 > /dev/null
for rep in 1 2; do
echo "== default u32+index 512x15x2 (rep $rep)"; python tools/quick_bench.py kinds=uniform,range perm=1 iters=10
for f in variants/librsb200_U32VA*.so; do echo "== $f (rep $rep)"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range perm=1 iters=10; done
echo "== default u64 512x11x2 (rep $rep)"; python tools/quick_bench.py kinds=uniform,range types=u64 iters=10
for f in variants/librsb200_U64A*.so; do echo "== $f (rep $rep)"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range types=u64 iters=10; done
echo "== default u64+index 512x9x2 (rep $rep)"; python tools/quick_bench.py kinds=uniform,range types=u64 perm=1 iters=10
for f in variants/librsb200_U64VA*.so; do echo "== $f (rep $rep)"; RSB200_LIB=$f python tools/quick_bench.py kinds=uniform,range types=u64 perm=1 iters=10; done
done
} > ${O}_quick.log 2>&1
grep -E "^==|Gkeys" ${O}_quick.log | cut -c1-200
bash tools/debug_checks.sh run r2h
