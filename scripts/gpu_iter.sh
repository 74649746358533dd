#!/bin/bash
# quick iteration: GPU tests, forward breakdown (DBGS), C2 + C4 bench lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1; grep -E "rc=|passed|failed|FAILED|^E  " gpurun_out/tests_summary.txt | head -30
rm -f gpurun_out/fwd_breakdown.txt
for d in ${DBGS:-0 3}; do
  HEBB_FWD_DBG=$d timeout 300 python scripts/fwd_breakdown.py >> gpurun_out/fwd_breakdown.txt 2>&1 || echo "dbg $d rc=$?" >> gpurun_out/fwd_breakdown.txt
done
cat gpurun_out/fwd_breakdown.txt
for w in c2 c4; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --layers-out gpurun_out/layers_$w.json > gpurun_out/bench_$w.json 2>gpurun_out/bench_$w.err || tail -5 gpurun_out/bench_$w.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('stages_ms'), d['roofline']['frac'])
PY
done
for extra in "" "--aten-backward"; do
  python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile $extra > gpurun_out/bench_c6$extra.json 2>gpurun_out/bench_c6.err || tail -5 gpurun_out/bench_c6.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c6$extra.json').read().strip().splitlines()[-1])
print('c6 $extra', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
done
