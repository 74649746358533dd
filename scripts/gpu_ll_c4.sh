#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --prec bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_ll_c4.log 2>&1; echo "launchlist rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c4.csv 45 > gpurun_out/launches_c4_summary.txt; head -50 gpurun_out/launches_c4_summary.txt
