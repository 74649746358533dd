#!/bin/bash
# 2-GPU box: the NCCL test (eager + recorded step), then the default bench line at N=1 and N=2 as the driver launches it.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest -q -p no:cacheprovider --timeout 380 -m gpu tests/test_gpu_parity.py -k "nccl or graph" > gpurun_out/t_graph3.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_graph3.log | cut -c1-300
B="--steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-layer-profile"
timeout 300 python bench.py --gpus 1 $B > gpurun_out/i_n1.json 2> gpurun_out/i_n1.err; echo "n1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/i_n2.json 2> gpurun_out/i_n2.err; echo "n2 rc=$?"
timeout 300 python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/i_c6.json 2> gpurun_out/i_c6.err; echo "c6 rc=$?"
timeout 300 python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile --graph > gpurun_out/i_c6g.json 2> gpurun_out/i_c6g.err; echo "c6 graph rc=$?"
for f in i_n1 i_n2 i_c6 i_c6g; do
  python - "$f" <<'PYEOF'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, 'ms', round(d['ms_per_step'], 4), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'], 'graph', d['config']['cuda_graph'], d['config']['optimizer'])
except Exception as e:
    print(f, 'FAILED', e)
    print(open(f'gpurun_out/{f}.err').read()[-1200:])
PYEOF
done
