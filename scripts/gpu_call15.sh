#!/bin/bash
# 8 GPUs: BASELINE configs[3] (dense crowds, data-parallel training with the NCCL all-reduce) and configs[4] (best-of-20 over the five splits, scene-sharded)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
run() { n=$1; tag=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n "$@" > $O/c15_${tag}_n$n.json 2> $O/c15_${tag}_n$n.err; echo "$tag n=$n rc=$?"; python -c "
import json; d=json.loads(open('$O/c15_${tag}_n$n.json').read().strip().splitlines()[-1]); print('$tag', $n, round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],2), 'ms', d.get('weights_identical_across_ranks'))"; grep -h "FAILED in stage" $O/c15_${tag}_n$n.err | head -2; }
run 8 train_c4_tc --mode train --scenes 128 --agents 256 --steps 5 --warmup 2 --train-gemm tc
run 4 train_c4_tc --mode train --scenes 128 --agents 256 --steps 5 --warmup 2 --train-gemm tc
run 8 c5 --config c5
run 8 c5_f32 --config c5 --prec f32
