#!/bin/bash
# the driver's scaling run on one 8-GPU node: reference arm + our arm at N = 1, 2, 4, 8 (its launch lines)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
for N in 1 2 4 8; do
  if [ $N = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N"; fi
  timeout 300 $L bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/scale_ref_n$N.json 2> $O/scale_ref_n$N.err; echo "ref N=$N rc=$?"
  timeout 600 $L bench.py --gpus $N --steps 20 --warmup 5 > $O/scale_n$N.json 2> $O/scale_n$N.err; echo "ours N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('$O/scale_n$N.json').read().strip().splitlines()[-1]); r=json.loads(open('$O/scale_ref_n$N.json').read().strip().splitlines()[-1])
    print('N=$N', round(d['value']/1e6,2), 'M/s', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e6,2), ' ref', round(r['value']/1e3,1), 'k/s  e2e ratio', round(d['e2e']['value']/r['value'],1), d['clocks'])
except Exception as e: print('N=$N no line', e)
PY
  grep -h "FAILED in stage" $O/scale_n$N.err | head -3
done
