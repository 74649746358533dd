#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests/test_umma_probe.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "rsw or run6 or run9" 2>&1 | tail -3
bash scripts/gpu_ll_c4.sh 2>&1 | head -14
