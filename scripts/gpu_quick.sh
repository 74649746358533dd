#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; grep -E "rc=|passed|failed|FAILED|^E  " gpurun_out/tests_summary.txt | head -30
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2>gpurun_out/bench_c2.err; tail -c 2500 gpurun_out/bench_c2.json
python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/bench_c1.json 2>gpurun_out/bench_c1.err; tail -c 1800 gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
