#!/bin/bash
# end of round, 2 GPUs: the driver's launch line (f16 default) and the packed training step (eager and CUDA graph) under NCCL
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602"
timeout 300 $L bench.py --gpus 2 --steps 20 --warmup 5 > $O/final2_bench.json 2> $O/final2_bench.err; echo "bench rc=$?"
timeout 200 $L bench.py --gpus 2 --mode train --variant mc --train-gemm tc --agents 256 --scenes 128 --steps 5 --warmup 3 > $O/final2_train_c4.json 2> $O/final2_train_c4.err; echo "train c4 rc=$?"
timeout 200 $L bench.py --gpus 2 --mode train --variant mcr --train-gemm tc --train-graph --scenes 512 --steps 5 --warmup 3 > $O/final2_train_mcr.json 2> $O/final2_train_mcr.err; echo "train mcr rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final2_bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value']/1e6,2), 'M/s', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,2), d['dtype'][:3], d['modes']['f16']['within_1e-3'])
for f in ('final2_train_c4','final2_train_mcr'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],2), 'ms identical weights', d['weights_identical_across_ranks'], d['loss_first'], d['loss_last'])
PY
