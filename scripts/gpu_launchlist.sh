#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "launchlist rc=$?"
tail -2 gpurun_out/plain.log | cut -c1-300
python scripts/profile_layer.py 128 64 3 64 64 64 bf16 > gpurun_out/pl3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_swta|fwd_swta' -s 6 -c 2 -o gpurun_out/prof_128x64_64_bf16 python scripts/profile_layer.py 128 64 3 64 64 64 bf16 > gpurun_out/ncu3.log 2>&1; echo "ncu3 rc=$?"
python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu tests/test_gpu_parity.py -k "tensor_core" 2>&1 | tail -3
