#!/bin/bash
# Run ON THE GPU BOX (via gpurun): plain bench run first, then the ncu launch list of the same command, then one
# `--set full` capture of the two top kernels.  Outputs land in gpurun_out/ (summarised into profiles/ by
# tools/summarize_profiles.py back in the build container).   usage: tools/gpu_profile.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-other"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/${TAG}_bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# 14 matching launches belong to the self-test sorts of rs_context_create, 5 to the first warm-up sort: capture the second warm-up sort (histogram + 4 passes)
ncu --set full --clock-control none --import-source on -k regex:"onesweep|histogram" -s 19 -c 5 -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
