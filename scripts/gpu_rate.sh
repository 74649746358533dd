#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python scripts/umma_rate.py > gpurun_out/umma_rate.txt 2>&1; echo "rate rc=$?"; cat gpurun_out/umma_rate.txt
