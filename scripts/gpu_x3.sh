#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python bench.py --workload c4 --prec bf16x3 --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_c4_bf16x3.json > gpurun_out/bench_c4_bf16x3.json 2>gpurun_out/bench_c4.err || tail -5 gpurun_out/bench_c4.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_c2_bf16x3.json > gpurun_out/bench_c2q.json 2>gpurun_out/bench_c2q.err || tail -5 gpurun_out/bench_c2q.err
python - <<PY
import json
for f,l in (('bench_c4_bf16x3','layers_c4_bf16x3'),('bench_c2q','layers_c2_bf16x3')):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('stage_ms'))
    for l in json.load(open('gpurun_out/%s.json'%l))['layers']:
        if l['Cout'] in (64,128) and l['Cin'] in (64,128) and l['k'][-1]==3:
            gf=l['flops_one_contraction']/1e9
            print(f"  {l['kind'][7:]:10s} {l['Cin']:4d}->{l['Cout']:4d} x{l['x'][2:]} fwd {l['fwd_ms']:.3f} dw {l['dw_ms']:.3f} ({gf/l['dw_ms']:.0f})")
PY
