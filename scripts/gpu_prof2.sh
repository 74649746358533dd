#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the two tcgen05 kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-layer-profile"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "launchlist rc=$?"
# tensor-bound layer of the 3-D network: 128->128 3x3x3 @48x48x40, batch 8
for P in bf16x3 bf16; do
python scripts/profile_layer.py 128 128 3 48 40 8 $P 48 > gpurun_out/pl_$P.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_swta|fwd_swta' -s 6 -c 2 -o gpurun_out/prof_c4_128x128_$P python scripts/profile_layer.py 128 128 3 48 40 8 $P 48 > gpurun_out/ncu_$P.log 2>&1; echo "ncu $P rc=$?"; cat gpurun_out/pl_$P.log
done
# HBM-bound layer of the 2-D network: 16->16 3x3 @256x256, batch 64
python scripts/profile_layer.py 16 16 3 256 256 64 bf16x3 > gpurun_out/pl_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw_swta|fwd_swta|pack_x' -s 9 -c 3 -o gpurun_out/prof_c2_16x16 python scripts/profile_layer.py 16 16 3 256 256 64 bf16x3 > gpurun_out/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
