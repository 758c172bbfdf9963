#!/bin/bash
# the driver's 2-GPU launch line with the f16 default (3 runs), and a 2-GPU mcr training step on the new edge kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602"
for i in 1 2 3; do
  timeout 600 $L bench.py --gpus 2 --steps 20 --warmup 5 > $O/scale2_f16_$i.json 2> $O/scale2_f16_$i.err; echo "run $i rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('$O/scale2_f16_$i.json').read().strip().splitlines()[-1])
    print(round(d['value']/1e6,2), 'M/s', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e6,2), d['dtype'][:4], d['modes']['f16']['within_1e-3'], d['clocks'])
except Exception as e: print('no line', e)
PY
done
timeout 600 $L bench.py --gpus 2 --mode train --variant mcr --train-gemm tc --train-graph --scenes 512 --steps 5 --warmup 3 > $O/train2_mcr_tc_graph.json 2> $O/train2.err; echo "train rc=$?"; tail -c 500 $O/train2_mcr_tc_graph.json
