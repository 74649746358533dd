#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; grep -E "rc=|passed|failed|FAILED|^E  " gpurun_out/tests_summary.txt | head -20
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-layer-profile "$@" 2>gpurun_out/err.log | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))" || tail -3 gpurun_out/err.log; }
run --no-fuse
run
run --head-channels-last
run --cudnn-benchmark
run --head-channels-last --cudnn-benchmark
run --workload c4
run --workload c4 --prec bf16
