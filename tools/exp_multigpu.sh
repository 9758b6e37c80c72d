#!/bin/bash
# Multi-GPU validation (run with gpurun --gpus N): the C++ host through rs_dist_create, exact parity of the native and
# the Python-orchestrated paths, the contract bench line, A/B of the merge walk order on range-partitioned input.
# usage: tools/exp_multigpu.sh <ngpus> [tag]
N=${1:-2}; TAG=${2:-r2g}
mkdir -p gpurun_out
O=gpurun_out/${TAG}_n${N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
./tests/cpp/dist_runner $N 22 > ${O}_dist_runner.txt 2>&1; echo "dist_runner rc=$?"; tail -4 ${O}_dist_runner.txt
$TR tools/dist_check.py exchange=native > ${O}_dist_check_native.txt 2>&1; echo "dist_check native rc=$?"; tail -2 ${O}_dist_check_native.txt
$TR bench.py --gpus $N --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench native rc=$?"; cut -c1-1500 ${O}_bench.json; tail -3 ${O}_bench.err
cp gpurun_out/dist_parity_n${N}.txt ${O}_bench_parity.txt 2>/dev/null
$TR bench.py --gpus $N --steps 5 --warmup 3 --exchange p2p > ${O}_bench_p2p.json 2> ${O}_bench_p2p.err; echo "bench p2p rc=$?"; cut -c1-400 ${O}_bench_p2p.json
RSB200_MERGE_ORDER_REGIONS=1 $TR bench.py --gpus $N --steps 5 --warmup 3 > ${O}_bench_regions.json 2> ${O}_bench_regions.err; echo "bench regions rc=$?"; python - <<PY
import json
for f in ("${O}_bench.json", "${O}_bench_regions.json", "${O}_bench_p2p.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "Gkeys/s", d["ms_per_step"], "ms", d["config"].get("datasets_gkeys_s"), d["roofline"].get("stage_ms_max_over_ranks"))
    except Exception as e:
        print(f, "unreadable", e)
PY
