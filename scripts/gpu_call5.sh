#!/bin/bash
# 8-GPU box: the driver's launch line at N = 4 and N = 8 (and N = 1, 2 for the curve), then the GPU test suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L | wc -l
for N in 8 4 2 1; do
  if [ $N = 1 ]; then cmd="python bench.py --gpus 1 --steps 20 --warmup 5 --no-modes"; else
  cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --no-modes"; fi
  timeout 400 $cmd > $O/c5_bench_n$N.json 2> $O/c5_bench_n$N.err; echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('$O/c5_bench_n$N.json').read().strip().splitlines()[-1])
    print('N=$N', round(d['value']/1e6,2), 'M/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']/1e6,2), 'clocks', d['clocks'])
except Exception as e: print('N=$N no line', e)
PY
  grep -h "FAILED in stage" $O/c5_bench_n$N.err | head -3
done
ls $O/rank*.err 2>/dev/null
echo "== nccl test"; timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
