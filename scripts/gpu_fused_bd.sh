#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for shape in "16 16 256 3" "3 16 256 3" "32 16 256 3" "32 32 128 3" "32 16 128 1"; do
  for f in 1 0; do HEBB_FUSED=$f timeout 120 python scripts/fused_breakdown.py $shape fwd 2>&1 | grep dbg= | cut -c1-70; done
done
HEBB_FUSED_PROF=1 timeout 120 python scripts/fused_breakdown.py 16 16 256 3 fwd 2>&1 | grep -v Warn | cut -c1-200
