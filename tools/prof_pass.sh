#!/bin/bash
# ncu --set full of ONE 2^28-key uint32 pass launch of the default and of the RSB200_PASS=3 kernel (after 12 self-test
# launches and 2 launches of the first warm-up sort).  usage: tools/prof_pass.sh <tag>
TAG=${1:-r2j}
mkdir -p gpurun_out
python tools/quick_bench.py kinds=uniform iters=3 > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"onesweep" -s 14 -c 1 -o gpurun_out/${TAG}_v1 -f python tools/quick_bench.py kinds=uniform iters=1 > gpurun_out/${TAG}_ncu_v1.log 2>&1; echo "v1 rc=$?"
RSB200_PASS=3 ncu --set full --clock-control none --import-source on -k regex:"pass3" -s 14 -c 1 -o gpurun_out/${TAG}_v3 -f python tools/quick_bench.py kinds=uniform iters=1 > gpurun_out/${TAG}_ncu_v3.log 2>&1; echo "v3 rc=$?"
ls -la gpurun_out/${TAG}_*
