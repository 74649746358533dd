#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --prec bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'swta|finalize|pack_|wnorm|softmax' -s 420 -c 140 --csv --log-file gpurun_out/launches_c4_bf16.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "rc=$?"
