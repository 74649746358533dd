#!/bin/bash
# Round-2 call C: all GPU tests, smoke, the default bench line (with companions), the same without the fused kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 900 -m gpu"
timeout 900 $PY tests/test_umma_probe.py > gpurun_out/t_probe.log 2>&1; echo "probe rc=$?"
timeout 1800 $PY tests/test_gpu_parity.py > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python bench.py --steps 10 --warmup 3 --layers-out gpurun_out/layers_c2.json > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
HEBB_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --layers-out gpurun_out/layers_c2_nofused.json > gpurun_out/bench_c2_nofused.json 2> gpurun_out/bench_c2_nofused.err; echo "bench(nofused) rc=$?"
timeout 300 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1 rc=$?"
timeout 300 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline --graph > gpurun_out/bench_c1_graph.json 2> gpurun_out/bench_c1_graph.err; echo "c1 graph rc=$?"
tail -n 30 gpurun_out/t_parity.log | cut -c1-300
tail -n 5 gpurun_out/smoke.log
tail -n 5 gpurun_out/bench_default.err
