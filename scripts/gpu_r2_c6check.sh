cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python bench.py --workload c6 --steps 5 --warmup 3 --no-cpu-baseline --no-layer-profile > gpurun_out/bench_c6.json 2> gpurun_out/bench_c6.err; echo "c6 rc=$?"; tail -3 gpurun_out/bench_c6.err | cut -c1-200
timeout 600 python -m pytest -q -p no:cacheprovider -m gpu -x tests/test_gpu_parity.py -k "wgrad or fused or f2d or graph or training_loop" 2>&1 | tail -2
for a in "3 16 256 3" "16 16 256 3" "32 16 256 3"; do timeout 100 python scripts/fused_breakdown.py $a 2>&1 | tail -1 | cut -c1-60; done
HEBB_FUSED_PROF=1 timeout 100 python scripts/fused_breakdown.py 3 16 256 3 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-layer-profile 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
python -c "
import json
d=json.loads(open('gpurun_out/bench_c6.json').read().strip().splitlines()[-1]); print('c6', d['ms_per_step'], d['value'])"
