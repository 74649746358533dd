#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for extra in "" "--head-wgrad 16" "--head-wgrad 32" "--head-wgrad 64"; do
  python bench.py --workload c2 --steps 5 --warmup 4 --no-cpu-baseline --no-layer-profile $extra > gpurun_out/bench_head.json 2>gpurun_out/bench_head.err || tail -5 gpurun_out/bench_head.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_head.json').read().strip().splitlines()[-1])
print('c2 [$extra]', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
done
