#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for d in ${DBGS:-0 3 19 35 67 64}; do
  HEBB_FWD_DBG=$d timeout 300 python scripts/fwd_breakdown.py >> gpurun_out/fwd_breakdown.txt 2>&1 || echo "dbg $d rc=$?" >> gpurun_out/fwd_breakdown.txt
done
cat gpurun_out/fwd_breakdown.txt
