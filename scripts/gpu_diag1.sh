#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "launchlist rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c2.csv 45 > gpurun_out/launches_c2_summary.txt; head -50 gpurun_out/launches_c2_summary.txt
i=0
for L in "128 64 3 96 80 8 bf16 96" "512 512 3 12 10 8 bf16 12" "64 64 3 96 80 8 bf16 96"; do
  i=$((i+1))
  python scripts/profile_layer.py $L > gpurun_out/pl_$i.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'dw_swta' -s 3 -c 1 -o gpurun_out/prof_dw_$i python scripts/profile_layer.py $L > gpurun_out/ncu_$i.log 2>&1; echo "ncu $i rc=$?"; cat gpurun_out/pl_$i.log
done
