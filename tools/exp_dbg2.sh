#!/bin/bash
# 2-GPU debugging of the level-1 merge: native path with page-rounded buffers, merge kernels timed; then the plain line
mkdir -p gpurun_out
O=gpurun_out/r2dbg2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
RSB200_MERGE_TIMING=1 $TR bench.py --gpus 2 --steps 2 --warmup 3 > ${O}_native_mt.json 2> ${O}_native_mt.err; echo "native rc=$?"; grep "rsb200 merge, device 0\] 10" ${O}_native_mt.err | tail -3
$TR bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_native.json 2> ${O}_native.err; echo "native plain rc=$?"
python - <<PY
import json
d = json.loads(open("${O}_native.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "Gkeys/s", round(d["ms_per_step"], 3), "ms", d["config"].get("datasets_gkeys_s"), d["roofline"].get("stage_ms_max_over_ranks"))
PY
