#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; grep -E "rc=|passed|failed|FAILED" gpurun_out/tests_summary.txt | head -20
python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_c4.json > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; tail -3 gpurun_out/bench_c4.err
python bench.py --workload c4 --prec bf16 --steps 2 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_c4_bf16.json > gpurun_out/bench_c4_bf16.json 2> gpurun_out/bench_c4_bf16.err; echo "bench c4 bf16 rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/bench_c4.json','gpurun_out/bench_c4_bf16.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), d['roofline']['stage_ms'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY
