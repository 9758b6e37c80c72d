#!/bin/bash
# Run ON THE GPU BOX (gpurun, 1 GPU): validates and times the opt-in paths that were written without GPU access.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'tools/validate_optins_1gpu.sh'
# Results -> gpurun_out/optins_1gpu.txt.  Build the tile-shape variants BEFORE the call (they travel with the snapshot):
#   EXTRA="-DRS_MERGE_THREADS=128 -DRS_MERGE_ITEMS=32 -DRS_MERGE_MINB32=6 -DRS_MERGE_MINB32_DENSE=6" tools/build_variants.sh "U32A 512 22 2"
mkdir -p gpurun_out
OUT=gpurun_out/optins_1gpu.txt
: > $OUT
echo "== merge tests, default walk order" >> $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k merge 2>&1 | tail -2 >> $OUT
echo "== merge tests, RSB200_MERGE_ORDER_REGIONS=1" >> $OUT
RSB200_MERGE_ORDER_REGIONS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k merge 2>&1 | tail -2 >> $OUT
echo "== torch entry tests" >> $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k torch_sort 2>&1 | tail -2 >> $OUT
echo "== local merge level 2^30 u32: default / regions order / 4 jobs" >> $OUT
python tools/merge_bench.py log2n=30 >> $OUT 2>&1
RSB200_MERGE_ORDER_REGIONS=1 python tools/merge_bench.py log2n=30 >> $OUT 2>&1
RSB200_MERGE_ORDER_REGIONS=1 python tools/merge_bench.py log2n=30 jobs=4 >> $OUT 2>&1
for lib in variants/*.so; do
  [ -e "$lib" ] || continue
  echo "== $lib" >> $OUT
  RSB200_LIB=$lib python tools/merge_bench.py log2n=30 >> $OUT 2>&1
done
cat $OUT
