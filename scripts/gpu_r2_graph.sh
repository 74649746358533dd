#!/bin/bash
# Round-2 call: CUDA-graph replay of the whole step (device-resident dropout state), fused Adam, tile-size rule.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 900 -m gpu"
timeout 900 $PY tests/test_gpu_parity.py -k "graph or dropout or fuse_pass or data_parallel" > gpurun_out/t_graph.log 2>&1; echo "graph tests rc=$?"
B="--steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-layer-profile"
timeout 300 python bench.py $B > gpurun_out/g_c2_eager.json 2> gpurun_out/g_c2_eager.err; echo "c2 eager rc=$?"
timeout 300 python bench.py $B --fused-adam > gpurun_out/g_c2_eager_fa.json 2> gpurun_out/g_c2_eager_fa.err; echo "c2 eager fused-adam rc=$?"
timeout 300 python bench.py $B --graph > gpurun_out/g_c2_graph.json 2> gpurun_out/g_c2_graph.err; echo "c2 graph rc=$?"
timeout 300 python bench.py $B --graph --fused-adam > gpurun_out/g_c2_graph_fa.json 2> gpurun_out/g_c2_graph_fa.err; echo "c2 graph fused-adam rc=$?"
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-layer-profile --graph > gpurun_out/g_c4_graph.json 2> gpurun_out/g_c4_graph.err; echo "c4 graph rc=$?"
timeout 300 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline --no-layer-profile --graph > gpurun_out/g_c1_graph.json 2> gpurun_out/g_c1_graph.err; echo "c1 graph rc=$?"
timeout 300 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline --no-layer-profile --graph --fused-adam > gpurun_out/g_c1_graph_fa.json 2> gpurun_out/g_c1_graph_fa.err; echo "c1 graph fa rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layers-out gpurun_out/g_layers_mb2.json > gpurun_out/g_c2_mb2.json 2> gpurun_out/g_c2_mb2.err; echo "c2 layers rc=$?"
HEBB_MB_WAVES=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layers-out gpurun_out/g_layers_mb1.json > gpurun_out/g_c2_mb1.json 2> gpurun_out/g_c2_mb1.err; echo "c2 layers mb1 rc=$?"
HEBB_MB_WAVES=0.9 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layers-out gpurun_out/g_layers_mb09.json > gpurun_out/g_c2_mb09.json 2> gpurun_out/g_c2_mb09.err; echo "c2 layers mb0.9 rc=$?"
tail -n 15 gpurun_out/t_graph.log | cut -c1-300
for f in g_c2_eager g_c2_eager_fa g_c2_graph g_c2_graph_fa g_c4_graph g_c1_graph g_c1_graph_fa g_c2_mb2 g_c2_mb1 g_c2_mb09; do
  python - "$f" <<'PYEOF'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, 'ms', round(d['ms_per_step'], 4), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'], 'graph', d['config']['cuda_graph'])
except Exception as e:
    print(f, 'FAILED', e)
    print(open(f'gpurun_out/{f}.err').read()[-1500:])
PYEOF
done
