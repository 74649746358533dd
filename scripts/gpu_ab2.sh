#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export HEBB_DW_REUSE=1
for W in "c4 bf16" "c2 bf16x3"; do
    set -- $W
    python bench.py --workload $1 --prec $2 --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_$1_1.json > gpurun_out/bench_$1_1.json 2>gpurun_out/bench_$1_1.err || tail -5 gpurun_out/bench_$1_1.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$1_1.json').read().strip().splitlines()[-1])
print('$1 reuse=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('stage_ms'), d['roofline']['frac'])
for l in json.load(open('gpurun_out/layers_$1_1.json'))['layers']:
    gf=l['flops_one_contraction']/1e9
    print(f"  {l['kind'][7:]:16s} {l['Cin']:5d}->{l['Cout']:5d} k{l['k'][-1]} x{l['x'][2:]} pack {l['pack_ms']:.3f} fwd {l['fwd_ms']:.3f} ({gf/l['fwd_ms']:.0f}) dw {l['dw_ms']:.3f} ({gf/l['dw_ms']:.0f})")
PY
done
