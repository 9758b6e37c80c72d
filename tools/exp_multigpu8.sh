#!/bin/bash
# 8-GPU validation, lean (8x the GPU-minutes): C++ runner, native bench line (all datasets + exact parity inside),
# Python-orchestrated path for A/B, one run with the merge kernels timed.
N=${1:-8}; TAG=${2:-r2g}
mkdir -p gpurun_out
O=gpurun_out/${TAG}_n${N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
./tests/cpp/dist_runner $N 22 > ${O}_dist_runner.txt 2>&1; echo "dist_runner rc=$?"; tail -3 ${O}_dist_runner.txt
$TR bench.py --gpus $N --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench native rc=$?"; tail -3 ${O}_bench.err
cp gpurun_out/dist_parity_n${N}.txt ${O}_bench_parity.txt 2>/dev/null
$TR bench.py --gpus $N --steps 5 --warmup 3 --exchange p2p > ${O}_bench_p2p.json 2> ${O}_bench_p2p.err; echo "bench p2p rc=$?"
RSB200_MERGE_TIMING=1 $TR bench.py --gpus $N --steps 2 --warmup 3 > ${O}_bench_mt.json 2> ${O}_bench_mt.err; echo "bench merge-timing rc=$?"; grep "rsb200 merge, device 0" ${O}_bench_mt.err | tail -6
RSB200_MERGE_ORDER_REGIONS=1 $TR bench.py --gpus $N --steps 5 --warmup 3 > ${O}_bench_regions.json 2> ${O}_bench_regions.err; echo "bench regions rc=$?"
python - <<PY
import json
for f in ("${O}_bench.json", "${O}_bench_p2p.json", "${O}_bench_regions.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "Gkeys/s", round(d["ms_per_step"], 3), "ms", d["config"].get("datasets_gkeys_s"), d["roofline"].get("stage_ms_max_over_ranks"), "e2e", d["e2e"].get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
