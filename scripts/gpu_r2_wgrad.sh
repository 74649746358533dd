#!/bin/bash
# hebb_conv_wgrad on the fused kernel: tests (each group in its own process: a trapping kernel poisons the context), bench with/without.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu -x"
timeout 600 $PY tests/test_gpu_parity.py -k "wgrad_on_the_fused" > gpurun_out/t_wg1.log 2>&1; echo "fused wgrad tests rc=$?"; tail -12 gpurun_out/t_wg1.log | cut -c1-250
python - <<'PYEOF'
import sys
sys.path.insert(0, 'hebbian-bootstraping-semi-supervised-medical-imaging_b200')
from hebb import _native
print('watchdog', _native.watchdog_code() if hasattr(_native, 'watchdog_code') else 'n/a')
PYEOF
timeout 600 $PY tests/test_gpu_parity.py -k "fast_wgrad or fuse_pass or fused_kernel_at_size or graph" > gpurun_out/t_wg2.log 2>&1; echo "related tests rc=$?"; tail -5 gpurun_out/t_wg2.log | cut -c1-250
B="--steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-layer-profile"
timeout 300 python bench.py $B > gpurun_out/w_c2.json 2> gpurun_out/w_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py $B --head-wgrad 16 > gpurun_out/w_c2_hw16.json 2> gpurun_out/w_c2_hw16.err; echo "c2 head-wgrad 16 rc=$?"
timeout 300 python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/w_c6.json 2> gpurun_out/w_c6.err; echo "c6 rc=$?"
HEBB_FUSED_WGRAD=0 timeout 300 python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/w_c6_off.json 2> gpurun_out/w_c6_off.err; echo "c6 off rc=$?"
HEBB_FUSED_WGRAD_PASSES=4 timeout 300 python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/w_c6_p4.json 2> gpurun_out/w_c6_p4.err; echo "c6 passes4 rc=$?"
for f in w_c2 w_c2_hw16 w_c6 w_c6_off w_c6_p4; do
  python - "$f" <<'PYEOF'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, 'ms', round(d['ms_per_step'], 4), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'], 'graph', d['config']['cuda_graph'])
except Exception as e:
    print(f, 'FAILED', e)
    print(open(f'gpurun_out/{f}.err').read()[-1200:])
PYEOF
done
