#!/bin/bash
# 2-GPU box: NCCL collectives recorded into the step's CUDA graph (test + bench lines), fused Adam.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu tests/test_gpu_parity.py -k "nccl or graph" > gpurun_out/t_graph2.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_graph2.log | cut -c1-300
B="--steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-layer-profile"
run1() { name=$1; shift; timeout 300 python bench.py --gpus 1 $B "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run2() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run1 h_n1_eager
run1 h_n1_eager_fa --fused-adam
run1 h_n1_graph --graph
run1 h_n1_graph_fa --graph --fused-adam
run2 h_n2_eager
run2 h_n2_eager_fa --fused-adam
run2 h_n2_graph --graph
run2 h_n2_graph_fa --graph --fused-adam
for f in h_n1_eager h_n1_eager_fa h_n1_graph h_n1_graph_fa h_n2_eager h_n2_eager_fa h_n2_graph h_n2_graph_fa; do
  python - "$f" <<'PYEOF'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, 'ms', round(d['ms_per_step'], 4), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'], 'graph', d['config']['cuda_graph'])
except Exception as e:
    print(f, 'FAILED', e)
    print(open(f'gpurun_out/{f}.err').read()[-1200:])
PYEOF
done
