#!/bin/bash
# 1 -> 8 GPU weak-scaling check of the C2 bench (and C4/bf16 at 8), as the driver launches it
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; echo "n1 rc=$?"
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29510+N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; echo "n$N rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --workload c4 --prec bf16 --steps 3 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/scale_c4_8.json 2> gpurun_out/scale_c4_8.err; echo "c4 n8 rc=$?"
python bench.py --gpus 1 --workload c4 --prec bf16 --steps 3 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/scale_c4_1.json 2> gpurun_out/scale_c4_1.err; echo "c4 n1 rc=$?"
tail -3 gpurun_out/scale_8.err
python - <<PY
import json
for n in ('1','2','4','8','c4_1','c4_8'):
    try:
        d=json.loads(open(f'gpurun_out/scale_{n}.json').read().strip().splitlines()[-1]); print(n, 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks'])
    except Exception as e: print(n, 'ERR', e)
PY
