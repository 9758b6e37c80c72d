#!/bin/bash
# compute-sanitizer over tools/sanitize_driver.py (run on the GPU box): memcheck, racecheck (shared-memory hazards: the
# ranking tables, the stage, the look-back hand-over words in shared memory), synccheck; both pass-kernel generations.
# Logs -> gpurun_out/<tag>_sanitizer_*.log (copied to profiles/ afterwards).   usage: tools/sanitize.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for v2 in 0 1; do
    log=gpurun_out/${TAG}_sanitizer_${tool}_sweepv$((v2+1)).log
    if [ $v2 = 1 ]; then export RSB200_SWEEP_V2=1; else unset RSB200_SWEEP_V2; fi
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py log2n=16 > $log 2>&1
    echo "$tool sweep_v$((v2+1)): rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver' $log | tr '\n' ' ')"
  done
done
unset RSB200_SWEEP_V2
