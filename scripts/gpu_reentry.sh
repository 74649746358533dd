#!/bin/bash
# re-entry check: GPU tests, smoke, C2 + C4(bf16) bench lines with per-layer tables
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; grep -E "rc=|passed|failed|FAILED|^E  " gpurun_out/tests_summary.txt | head -30
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_c2.json > gpurun_out/bench_c2.json 2>gpurun_out/bench_c2.err || tail -5 gpurun_out/bench_c2.err
tail -c 3000 gpurun_out/bench_c2.json
python bench.py --workload c4 --prec bf16 --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_c4_bf16.json > gpurun_out/bench_c4_bf16.json 2>gpurun_out/bench_c4.err || tail -5 gpurun_out/bench_c4.err
tail -c 2500 gpurun_out/bench_c4_bf16.json
