#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for extra in "" "--no-head-wgrad"; do
  python bench.py --workload c2 --steps 5 --warmup 4 --no-cpu-baseline --no-layer-profile $extra > gpurun_out/bench_hw.json 2>gpurun_out/bench_hw.err || tail -5 gpurun_out/bench_hw.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_hw.json').read().strip().splitlines()[-1])
print('c2 [$extra]', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
done
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer-profile"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_hw.csv $CMD > gpurun_out/ncu_ll_hw.log 2>&1; echo "launchlist rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c2_hw.csv 40
