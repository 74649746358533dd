#!/bin/bash
# usage: gpu_multi_r2.sh N   -- the 2-rank NCCL test, then the weak-scaling bench at 1 .. N GPUs as the driver launches it
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
timeout 600 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_parity.py -k "nccl or follow_their_tensors" > gpurun_out/t_nccl.log 2>&1; echo "nccl test rc=$?"; tail -3 gpurun_out/t_nccl.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-layer-profile > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; echo "N=1 rc=$?"
for n in 2 4 8; do
  if [ $n -le $N ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-layer-profile > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; echo "N=$n rc=$?"
  fi
done
for n in 1 2 4 8; do [ -f gpurun_out/scale_$n.json ] && python - gpurun_out/scale_$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), d['clocks']['reasons'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
