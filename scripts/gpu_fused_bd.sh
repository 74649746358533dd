#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for i in 0 1 2; do timeout 60 python scripts/fused_debug.py $i 2>&1 | grep -v "Warning\|detach\|float(" | grep -v "^case"; done
HEBB_FUSED_PROF=1 timeout 120 python scripts/fused_breakdown.py 3 16 256 3 2>&1 | grep -v Warn | cut -c1-200
timeout 300 python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu tests/test_gpu_parity.py -k "gather or at_size" 2>&1 | tail -2
