#!/bin/bash
# Run ON THE GPU BOX (gpurun --gpus 2): the opt-in multi-GPU paths against the defaults.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'tools/validate_optins_2gpu.sh'
# Results -> gpurun_out/optins_2gpu.txt
mkdir -p gpurun_out
OUT=gpurun_out/optins_2gpu.txt
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for env in "" "RSB200_DEVICE_SPLITTERS=1" "RSB200_MERGE_ORDER_REGIONS=1" "RSB200_DEVICE_SPLITTERS=1 RSB200_MERGE_ORDER_REGIONS=1"; do
  echo "== [$env] dist_check" >> $OUT
  env $env timeout 200 $TR tools/dist_check.py log2n_total=23 2>&1 | grep -E "PASSED|FAIL|Error|error" | tail -3 >> $OUT
  echo "== [$env] bench --gpus 2 (uniform), stages of the last step" >> $OUT
  env $env RSB200_DIST_TIMING=1 timeout 300 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/optins_2gpu_bench.log 2>&1
  grep -oE "\[rank [0-9]\] dist timing[^\[]*" gpurun_out/optins_2gpu_bench.log | tail -2 >> $OUT
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/optins_2gpu_bench.log | head -1 >> $OUT
done
echo "== range dataset (every rank keeps its own keys: all tiles of level 1 in one job), default vs regions order" >> $OUT
for env in "" "RSB200_MERGE_ORDER_REGIONS=1"; do
  env $env RSB200_DIST_TIMING=1 timeout 300 $TR bench.py --gpus 2 --steps 3 --warmup 3 --dataset range > gpurun_out/optins_2gpu_bench.log 2>&1
  echo "[$env]" >> $OUT
  grep -oE "\[rank [0-9]\] dist timing[^\[]*" gpurun_out/optins_2gpu_bench.log | tail -2 >> $OUT
done
cat $OUT
