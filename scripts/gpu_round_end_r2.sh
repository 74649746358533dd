#!/bin/bash
# Round-2 evidence for profiles/: the driver's own test command, smoke, both bench arms (the default line carries the C4
# sub-lines, the plain drop-in, the GPU-eager reference and the element-wise kernels), the eagerly launched variants, C1, C5, C6,
# the ncu launch list and DRAM-traffic capture of the bench command, full ncu captures of the fused kernel and of the two
# tcgen05 kernels on a tensor-bound 3-D layer.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --layers-out gpurun_out/layers_c2_bf16x3.json > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
HEBB_FUSED=0 python bench.py --no-cpu-baseline --no-extras --layers-out gpurun_out/layers_c2_bf16x3_nofused.json > gpurun_out/bench_c2_nofused.json 2> gpurun_out/bench_c2_nofused.err; echo "bench nofused rc=$?"
python bench.py --no-cpu-baseline --no-extras --no-layer-profile --no-graph --no-fused-adam --head-wgrad 16 > gpurun_out/bench_c2_eager.json 2> gpurun_out/bench_c2_eager.err; echo "bench eager (no graph, for-each Adam, cuDNN head wgrad) rc=$?"
python bench.py --workload c1 --steps 50 --warmup 5 > gpurun_out/bench_c1_graph.json 2> gpurun_out/bench_c1_graph.err; echo "c1 (graph) rc=$?"
python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline --no-graph --no-fused-adam > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1 eager rc=$?"
python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/bench_c6.json 2> gpurun_out/bench_c6.err; echo "c6 rc=$?"
python bench.py --workload c5 --steps 10 --warmup 3 --no-layer-profile > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
# (the profiler passes launch the step eagerly: the same kernels as the replayed graph, a known launch count per step)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-layer-profile --no-extras --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "launchlist rc=$?"
# DRAM traffic of our kernels over the whole run (7 steps: 3 warm-up, 2 timed, 2 end-to-end)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'swta|pack_x|pack_w|tc_finalize|wnorm_kernel|fused_small|fused_prep|fused_wgrad' --csv --log-file gpurun_out/traffic_c2.csv $CMD > gpurun_out/ncu_tr.log 2>&1; echo "traffic rc=$?"
# the fused small-channel kernel: 16->16 3x3 @256x256, batch 64
python scripts/fused_breakdown.py 16 16 256 3 > gpurun_out/pl_fused.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_small_kernel -s 2 -c 1 -f -o gpurun_out/r2_fused_16x16 python scripts/fused_breakdown.py 16 16 256 3 > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused rc=$?"; cat gpurun_out/pl_fused.log | cut -c1-80
# tensor-bound layer of the 3-D network: 128->128 3x3x3 @48x48x40, batch 8
python scripts/profile_layer.py 128 128 3 48 40 8 bf16 48 > gpurun_out/pl_bf16.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_swta|fwd_swta' -s 6 -c 2 -f -o gpurun_out/r2_c4_128x128_bf16 python scripts/profile_layer.py 128 128 3 48 40 8 bf16 48 > gpurun_out/ncu_bf16.log 2>&1; echo "ncu c4 rc=$?"; cat gpurun_out/pl_bf16.log
# hebb_conv_wgrad on the fused kernel against cuDNN on the 2-D head's layers, and its wait-cycle table (4-converter-warp build)
python scripts/wgrad_bench.py 64 > gpurun_out/wgrad_bench.txt 2>&1; echo "wgrad bench rc=$?"; tail -5 gpurun_out/wgrad_bench.txt | cut -c1-160
HEBB_FUSED_PROF=1 HEBB_FUSED_WG_NCW=4 python scripts/wgrad_breakdown.py 32 32 nhwc > gpurun_out/wgrad_prof.txt 2>&1; echo "wgrad prof rc=$?"
for f in gpurun_out/bench_c2.json gpurun_out/bench_c2_nofused.json gpurun_out/bench_c2_eager.json gpurun_out/bench_c6.json gpurun_out/bench_c1.json gpurun_out/bench_c1_graph.json gpurun_out/bench_c5.json gpurun_out/bench_ref.json; do
python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline') or {}
    print(sys.argv[1], d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), r.get('stage_ms'), r.get('frac'), (d.get('clocks') or {}).get('reasons'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
