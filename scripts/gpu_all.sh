#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; head -5 gpurun_out/tests_summary.txt; grep -E "passed|failed" gpurun_out/tests_summary.txt
python scripts/umma_rate.py > gpurun_out/umma_rate.txt 2>&1; echo "rate rc=$?"; grep "ctas=148" gpurun_out/umma_rate.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_c2.json > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --prec bf16 --no-cpu-baseline --layers-out gpurun_out/layers_c2_bf16.json > gpurun_out/bench_c2_bf16.json 2> gpurun_out/bench_c2_bf16.err; echo "bench bf16 rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/bench_c2.json','gpurun_out/bench_c2_bf16.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['roofline']['stage_ms'] if d['roofline'] else None)
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench_c2.err
