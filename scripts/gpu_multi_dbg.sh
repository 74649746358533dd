#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export HEBB_BENCH_E2E_DEBUG=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/scale_2.json 2> gpurun_out/scale_2.err; echo "n2 rc=$?"
grep "e2e rank" gpurun_out/scale_2.err
OMP_NUM_THREADS=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/scale_2b.json 2> gpurun_out/scale_2b.err; echo "n2b rc=$?"
grep "e2e rank" gpurun_out/scale_2b.err
python - <<PY
import json
for n in ('2','2b'):
    d=json.loads(open(f'gpurun_out/scale_{n}.json').read().strip().splitlines()[-1]); print(n, 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['e2e']['ms_per_step'])
PY
