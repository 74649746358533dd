#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 0 1 2 3 4 5 6 7 8; do timeout 60 python scripts/fused_debug.py $i 2>&1 | grep -v "Warning\|detach\|float(" | grep -v "^case" ; done
for shape in "16 16 256 3" "32 16 256 3" "16 32 128 3" "32 16 128 1"; do
  for m in 0 7; do HEBB_FUSED_PROF=1 HEBB_FUSED_DBG=$m timeout 120 python scripts/fused_breakdown.py $shape 2>&1 | grep -v Warn | cut -c1-200; done
done
timeout 900 python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu tests/test_gpu_parity.py > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/t_parity.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --layers-out gpurun_out/layers_c2.json > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; grep -E "Error" gpurun_out/bench_c2.err | tail -3
