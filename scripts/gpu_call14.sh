#!/bin/bash
# 2 GPUs: data-parallel training (synthetic C3-shaped, C4 dense crowds, real-data c2) with the NCCL gradient all-reduce
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 "$@" > $O/c14_$tag.json 2> $O/c14_$tag.err; echo "$tag rc=$?"; python -c "
import json; d=json.loads(open('$O/c14_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],2), 'ms', d.get('loss_first'), d.get('loss_last'), d.get('weights_identical_across_ranks'), d.get('held_out_zara01_best_of_20'))"; grep -h "FAILED in stage" $O/c14_$tag.err | head -2; }
run train_c3_tc --mode train --scenes 1024 --steps 5 --warmup 2 --train-gemm tc
run train_c4_tc --mode train --scenes 128 --agents 256 --steps 5 --warmup 2 --train-gemm tc
run train_mcr_tc --mode train --variant mcr --scenes 512 --steps 3 --warmup 1 --train-gemm tc
run c2 --config c2 --steps 10 --warmup 2 --train-gemm tc
run c5 --config c5
timeout 300 python bench.py --mode train --scenes 128 --agents 256 --steps 5 --warmup 2 --train-gemm tc > $O/c14_train_c4_tc_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('$O/c14_train_c4_tc_n1.json').read().strip().splitlines()[-1]); print('c4 n1', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],2))"
