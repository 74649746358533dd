cd "${GRAFT_REPO_ROOT:-/root/repo}"
L=hebbian-bootstraping-semi-supervised-medical-imaging_b200/hebb
for v in base own32; do
  echo "== $v"
  if [ $v = own32 ]; then cp $L/libhebb_sm100.so /tmp/base.so; cp $L/libhebb_sm100_own32.so $L/libhebb_sm100.so; fi
  for a in "32 16 256 3" "16 32 128 3" "16 16 256 3"; do timeout 100 python scripts/fused_breakdown.py $a 2>&1 | tail -1 | cut -c1-60; done
  timeout 300 python -m pytest -q -p no:cacheprovider -m gpu -x tests/test_gpu_parity.py -k "fused_kernel_at_size or f2d" 2>&1 | tail -1
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-layer-profile 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
