#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.json
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; grep -E "rc=|passed|failed|FAILED|^E  " gpurun_out/tests_summary.txt | head -20
for W in c2 c4; do for P in bf16x3 bf16; do
python bench.py --workload $W --prec $P --steps 3 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_${W}_${P}.json > gpurun_out/bench_${W}_${P}.json 2> gpurun_out/bench_${W}_${P}.err; echo "bench $W $P rc=$?"
done; done
python - <<'PY'
import json
for w in ['c2','c4']:
  for p in ['bf16x3','bf16']:
    f=f'gpurun_out/bench_{w}_{p}.json'
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()}, round(d['roofline']['frac'],4))
    except Exception as e: print(f, 'ERR', e)
PY
