#!/bin/bash
# Round-2 call B: the fused small-channel kernel -- its own tests first (a trap poisons the context), then everything.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu -x"
timeout 600 $PY tests/test_gpu_parity.py -k "f2d_ and bf16x3" > gpurun_out/t_fused1.log 2>&1; echo "fused-small rc=$?"
timeout 600 $PY tests/test_gpu_parity.py -k "fused_kernel_at_size or c2d_16_16 or c2d_32_32 or additivity or sum_to_one or batchnorm_statistics" > gpurun_out/t_fused2.log 2>&1; echo "fused-size rc=$?"
timeout 1500 python -m pytest -q -p no:cacheprovider --timeout 900 -m gpu tests/test_gpu_parity.py > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/bench_c2_fused.json 2> gpurun_out/bench_c2_fused.err; echo "bench rc=$?"
HEBB_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/bench_c2_nofused.json 2> gpurun_out/bench_c2_nofused.err; echo "bench(nofused) rc=$?"
tail -n 30 gpurun_out/t_fused1.log; tail -n 30 gpurun_out/t_fused2.log; tail -n 15 gpurun_out/t_parity.log
cut -c1-400 gpurun_out/bench_c2_fused.json; echo; cut -c1-400 gpurun_out/bench_c2_nofused.json
