#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; grep -E "rc=|passed|failed|FAILED|^E  " gpurun_out/tests_summary.txt | head -30
i=0
for L in "64 64 3 96 80 8 bf16 96" "128 64 3 96 80 8 bf16 96"; do
  i=$((i+1))
  python scripts/profile_layer.py $L > gpurun_out/pl_$i.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'dw_swta' -s 3 -c 1 -o gpurun_out/prof_dw2_$i python scripts/profile_layer.py $L > gpurun_out/ncu_$i.log 2>&1; echo "ncu $i rc=$?"; cat gpurun_out/pl_$i.log
done
