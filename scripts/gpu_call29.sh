#!/bin/bash
# edge-MLP backward kernels: parity tests, then the g2k_lstm_mcr training step (tc and fp32) at 512 scenes x 64 agents
mkdir -p gpurun_out
MMT_RECORD_ERRORS=gpurun_out/edge_bwd_errors.jsonl timeout 600 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "edge_mlp_backward or attention_score_grad or train_gradients or edge_mlp" > gpurun_out/edge_bwd_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/edge_bwd_tests.log
tail -15 gpurun_out/edge_bwd_tests.log
for g in tc fp32; do
  timeout 300 python bench.py --mode train --variant mcr --train-gemm $g --scenes 512 --steps 3 --warmup 3 \
    > gpurun_out/train_mcr_$g.json 2> gpurun_out/train_mcr_$g.err
  echo "train mcr $g rc=$?"; tail -c 600 gpurun_out/train_mcr_$g.json; tail -3 gpurun_out/train_mcr_$g.err
done
timeout 300 python bench.py --mode train --variant mc --train-gemm tc --scenes 1024 --steps 5 --warmup 3 > gpurun_out/train_mc_tc.json 2>/dev/null
tail -c 400 gpurun_out/train_mc_tc.json
