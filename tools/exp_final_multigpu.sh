#!/bin/bash
# Final multi-GPU validation of a round (run with gpurun --gpus N): the C++ host through rs_dist_create (exact, memcmp),
# the contract bench line (exact reduced-size parity + all five datasets inside), one run with the merge kernels timed.
# usage: tools/exp_final_multigpu.sh <ngpus> [tag]
N=${1:-2}; TAG=${2:-r2z}
mkdir -p gpurun_out
O=gpurun_out/${TAG}_n${N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
./tests/cpp/dist_runner $N 22 > ${O}_dist_runner.txt 2>&1; echo "dist_runner rc=$?"; tail -2 ${O}_dist_runner.txt
$TR bench.py --gpus $N --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench native rc=$?"; tail -3 ${O}_bench.err
cp gpurun_out/dist_parity_n${N}.txt ${O}_bench_parity.txt 2>/dev/null
RSB200_MERGE_TIMING=1 $TR bench.py --gpus $N --steps 2 --warmup 3 > ${O}_bench_mt.json 2> ${O}_bench_mt.err; echo "bench merge-timing rc=$?"
grep "tiles:" ${O}_bench_mt.err | awk '$5 > 100000' | tail -$((4 * N)) > ${O}_merge_timing.txt; tail -$N ${O}_merge_timing.txt
python - <<PY
import json
for f in ("${O}_bench.json", "${O}_bench_mt.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "Gkeys/s", round(d["ms_per_step"], 3), "ms", d["config"].get("datasets_gkeys_s"), d["roofline"].get("stage_ms_max_over_ranks"), "e2e", d["e2e"].get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
