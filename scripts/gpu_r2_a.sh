#!/bin/bash
# Round-2 call A: all GPU tests (separate processes per group), smoke, default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
PY="python -m pytest -q -p no:cacheprovider --timeout 900 -m gpu"
timeout 900 $PY tests/test_umma_probe.py > gpurun_out/t_probe.log 2>&1; echo "probe rc=$?"
timeout 1800 $PY tests/test_gpu_parity.py > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
tail -n 40 gpurun_out/t_probe.log gpurun_out/t_parity.log
tail -n 8 gpurun_out/smoke.log
head -c 1500 gpurun_out/bench_c2.json
