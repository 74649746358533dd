#!/bin/bash
# full kernel launch list of one C4 step (all kernels, ours and stock), for the step breakdown
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_c4_all.csv $CMD > gpurun_out/ncu_ll_c4.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c4_all.csv 45
