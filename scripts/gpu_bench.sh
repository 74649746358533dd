#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu tests/test_gpu_parity.py -k "network or additivity or stepper" > gpurun_out/t_net.log 2>&1; echo "net rc=$?"; tail -5 gpurun_out/t_net.log
python bench.py --steps 5 --warmup 3 --layers-out gpurun_out/layers_c2.json > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
python bench.py --steps 5 --warmup 3 --prec bf16 --no-cpu-baseline --layers-out gpurun_out/layers_c2_bf16.json > gpurun_out/bench_c2_bf16.json 2> gpurun_out/bench_c2_bf16.err; echo "bench bf16 rc=$?"
tail -c 1500 gpurun_out/bench_c2_bf16.json
