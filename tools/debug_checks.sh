#!/bin/bash
# Our own bounds checks in place of compute-sanitizer (closed on this GPU pool): run ON THE GPU BOX against a library built
# with -DRS_DEBUG_CHECKS (variants/librsb200_debugchecks.so, built by `tools/debug_checks.sh build` in the container):
# every scatter position, output address and look-back sum of both pass-kernel generations and of the one-CTA sort is
# checked on the device (RS_DCHECK traps), on the parity workloads of tools/sanitize_driver.py and tools/debug_chunks.py.
# usage: tools/debug_checks.sh build | tools/debug_checks.sh run <tag>
set -u
cd "$(dirname "$0")/.."
if [ "${1:-}" = build ]; then
  mkdir -p variants
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 177 -DRS_DEBUG_CHECKS \
       -shared -o variants/librsb200_debugchecks.so radix-sort_b200/csrc/rs_api.cu && echo built variants/librsb200_debugchecks.so
  exit $?
fi
TAG=${2:-r2}
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}_debug_checks.txt
export RSB200_LIB=variants/librsb200_debugchecks.so
{
echo "# library built with -DRS_DEBUG_CHECKS; any failed check prints RS_DCHECK and aborts the kernel (the run then fails)"
for v2 in 0 1; do
  if [ $v2 = 1 ]; then export RSB200_SWEEP_V2=1; else unset RSB200_SWEEP_V2; fi
  echo "== pass kernel generation $((v2+1)): sanitize_driver log2n=20"
  python tools/sanitize_driver.py log2n=20 2>&1 | tail -3; echo "rc=${PIPESTATUS[0]}"
  echo "== pass kernel generation $((v2+1)): chunk hand-over (RSB200_DEBUG_TILES_PER_CHUNK=3)"
  RSB200_DEBUG_TILES_PER_CHUNK=3 python tools/debug_chunks.py 2>&1 | grep -c " OK " ; echo "of 24 OK"
done
unset RSB200_SWEEP_V2
echo "== full size 2^28 u32 keys + index, uniform, through the checked build"
python tools/quick_bench.py kinds=uniform perm=1 iters=1 2>&1 | tail -1
} > $LOG 2>&1
grep -c RS_DCHECK $LOG; tail -12 $LOG
