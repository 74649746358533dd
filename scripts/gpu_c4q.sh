#!/bin/bash
# quick: GPU tests + C4 bf16 bench with per-layer table
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; grep -E "rc=|passed|failed|FAILED|^E  " gpurun_out/tests_summary.txt | head -30
python bench.py --workload c4 --prec bf16 --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_c4_bf16.json > gpurun_out/bench_c4_bf16.json 2>gpurun_out/bench_c4.err || tail -5 gpurun_out/bench_c4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c4_bf16.json').read().strip().splitlines()[-1])
print('c4', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('stage_ms'), d['roofline']['frac'])
for l in json.load(open('gpurun_out/layers_c4_bf16.json'))['layers']:
    gf=l['flops_one_contraction']/1e9
    print(f"{l['kind'][7:]:16s} {l['Cin']:5d}->{l['Cout']:5d} x{l['x'][2:]} pack {l['pack_ms']:.3f} fwd {l['fwd_ms']:.3f} ({gf/l['fwd_ms']:.0f}) dw {l['dw_ms']:.3f} ({gf/l['dw_ms']:.0f})")
PY
