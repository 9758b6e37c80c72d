#!/bin/bash
# Final single-GPU validation of a round: the whole GPU suite, smoke, the size sweep, the contract bench, the profiles.
# usage: tools/exp_final_n1.sh <tag>
TAG=${1:-r2z}
mkdir -p gpurun_out
O=gpurun_out/${TAG}
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
python tools/sweep.py > ${O}_size_sweep_u32.txt 2>&1; echo "sweep rc=$?"; tail -2 ${O}_size_sweep_u32.txt
python bench.py > ${O}_bench_n1.json 2> ${O}_bench_n1.err; echo "bench rc=$?"; cut -c1-200 ${O}_bench_n1.json
bash tools/gpu_profile.sh ${TAG}
