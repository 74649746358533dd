#!/bin/bash
# Round-end evidence: the driver's own test command, smoke, bench lines, launch list, DRAM traffic of our kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.json
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --layers-out gpurun_out/layers_c2_bf16x3.json > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "launchlist rc=$?"
# DRAM traffic of our kernels over one timed step (3 warm-up steps x 110 matching launches skipped)
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'swta|pack_x|pack_w|tc_finalize|wnorm_kernel' -s 396 -c 132 --csv --log-file gpurun_out/traffic_c2.csv $CMD > gpurun_out/ncu_tr.log 2>&1; echo "traffic rc=$?"
python bench.py --workload c5 --steps 10 --warmup 3 --no-layer-profile > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"; cut -c1-400 gpurun_out/bench_c5.json
