#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python scripts/umma_rate.py > gpurun_out/umma_rate.txt 2>&1; echo "rate rc=$?"; cat gpurun_out/umma_rate.txt
python scripts/profile_layer.py 16 16 3 256 256 64 bf16x3 > gpurun_out/pl.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_swta|fwd_swta' -s 6 -c 2 -o gpurun_out/prof_16x16_256 python scripts/profile_layer.py 16 16 3 256 256 64 bf16x3 > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc=$?"; tail -3 gpurun_out/ncu1.log
python scripts/profile_layer.py 64 64 3 64 64 64 bf16x3 > gpurun_out/pl2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_swta|fwd_swta' -s 6 -c 2 -o gpurun_out/prof_64x64_64 python scripts/profile_layer.py 64 64 3 64 64 64 bf16x3 > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"; tail -3 gpurun_out/ncu2.log
