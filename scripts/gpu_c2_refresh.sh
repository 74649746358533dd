#!/bin/bash
# refresh the C2 evidence only: bench lines (bf16x3 with the CPU arm, bf16), launch list, DRAM traffic window
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python bench.py --layers-out gpurun_out/layers_c2_bf16x3.json > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python bench.py --prec bf16 --no-cpu-baseline --layers-out gpurun_out/layers_c2_bf16.json > gpurun_out/bench_c2_bf16.json 2> gpurun_out/bench_c2_bf16.err; echo "bench bf16 rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "launchlist rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'swta|pack_x|pack_w|tc_finalize|wnorm_kernel' -s 396 -c 132 --csv --log-file gpurun_out/traffic_c2.csv $CMD > gpurun_out/ncu_tr.log 2>&1; echo "traffic rc=$?"
for f in gpurun_out/bench_c2.json gpurun_out/bench_c2_bf16.json; do cut -c1-200 $f; done
