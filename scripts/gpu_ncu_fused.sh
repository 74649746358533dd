#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python scripts/fused_breakdown.py 16 16 256 3 > gpurun_out/plain_fused.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_small_kernel -s 2 -c 1 -f -o gpurun_out/r2_fused_16x16 python scripts/fused_breakdown.py 16 16 256 3 > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fused.log; ls -la gpurun_out/*.ncu-rep
