#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "launchlist rc=$?"
python bench.py --steps 5 --warmup 3 --layers-out gpurun_out/layers_c2_bf16x3.json > gpurun_out/bench_c2.json 2>gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/bench_c2.json
