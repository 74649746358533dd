#!/bin/bash
# Round-end evidence for profiles/: the driver's own test command, smoke, both bench arms, C4/C5/C6 lines,
# the ncu launch list and DRAM-traffic capture of the bench command, full ncu captures of the two tcgen05 kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --layers-out gpurun_out/layers_c2_bf16x3.json > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python bench.py --prec bf16 --no-cpu-baseline --layers-out gpurun_out/layers_c2_bf16.json > gpurun_out/bench_c2_bf16.json 2> gpurun_out/bench_c2_bf16.err; echo "bench bf16 rc=$?"
for P in bf16x3 bf16; do
  python bench.py --workload c4 --prec $P --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_c4_$P.json > gpurun_out/bench_c4_$P.json 2> gpurun_out/bench_c4_$P.err; echo "c4 $P rc=$?"
done
python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1 rc=$?"
python bench.py --workload c5 --steps 10 --warmup 3 --no-layer-profile > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/bench_c6.json 2> gpurun_out/bench_c6.err; echo "c6 rc=$?"
python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile --aten-backward > gpurun_out/bench_c6_aten.json 2> gpurun_out/bench_c6_aten.err; echo "c6 aten rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "launchlist rc=$?"
# the same launch list for the 3-D network (bf16 operands)
CMD4="python bench.py --workload c4 --prec bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD4 > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv $CMD4 > gpurun_out/ncu_ll_c4.log 2>&1; echo "launchlist c4 rc=$?"
# DRAM traffic of our kernels over one timed step (3 warm-up steps x 132 matching launches skipped)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'swta|pack_x|pack_w|tc_finalize|wnorm_kernel' -s 396 -c 132 --csv --log-file gpurun_out/traffic_c2.csv $CMD > gpurun_out/ncu_tr.log 2>&1; echo "traffic rc=$?"
# tensor-bound layer of the 3-D network: 128->128 3x3x3 @48x48x40, batch 8
for P in bf16x3 bf16; do
python scripts/profile_layer.py 128 128 3 48 40 8 $P 48 > gpurun_out/pl_$P.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_swta|fwd_swta' -s 6 -c 2 -o gpurun_out/prof_c4_128x128_$P python scripts/profile_layer.py 128 128 3 48 40 8 $P 48 > gpurun_out/ncu_$P.log 2>&1; echo "ncu $P rc=$?"; cat gpurun_out/pl_$P.log
done
# shared-memory-bound layer of the 3-D network: 64->64 3x3x3 @96x96x80, batch 8, bf16 (k-step-outer loop, A collector re-use)
python scripts/profile_layer.py 64 64 3 96 80 8 bf16 96 > gpurun_out/pl_c4_64.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_swta' -s 3 -c 1 -o gpurun_out/prof_c4_64x64_bf16 python scripts/profile_layer.py 64 64 3 96 80 8 bf16 96 > gpurun_out/ncu_c4_64.log 2>&1; echo "ncu c4 64 rc=$?"
python scripts/umma_rate3.py > gpurun_out/umma_rate3.txt 2>&1; echo "rate3 rc=$?"
# small-channel layer of the 2-D network: 16->16 3x3 @256x256, batch 64
python scripts/profile_layer.py 16 16 3 256 256 64 bf16x3 > gpurun_out/pl_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_swta|fwd_swta|pack_x' -s 9 -c 3 -o gpurun_out/prof_c2_16x16 python scripts/profile_layer.py 16 16 3 256 256 64 bf16x3 > gpurun_out/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
for f in gpurun_out/bench_c2.json gpurun_out/bench_c2_bf16.json gpurun_out/bench_c4_bf16x3.json gpurun_out/bench_c4_bf16.json gpurun_out/bench_c1.json gpurun_out/bench_c5.json gpurun_out/bench_c6.json gpurun_out/bench_c6_aten.json gpurun_out/bench_ref.json; do
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
