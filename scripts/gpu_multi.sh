#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; echo "n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; echo "n$N rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/scale_ref_$N.json 2> gpurun_out/scale_ref_$N.err; echo "ref n$N rc=$?"; cut -c1-200 gpurun_out/scale_ref_$N.json
tail -3 gpurun_out/scale_$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.loads(open(f'gpurun_out/scale_{n}.json').read().strip().splitlines()[-1]); print(n, 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(n, 'ERR', e)
PY
