#!/bin/bash
# ncu --set full of the kernels added late in round 2: the f16 instantiation of the fused rollout, the edge-MLP backward
# (tcgen05) and the attention-logit gradient
mkdir -p gpurun_out
timeout 300 python scratch/step_profile.py f16 train-mcr > gpurun_out/step_profile_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"rollout_tc_kernel|edge_mlp_bwd_tc_kernel|attention_score_grad_kernel|edge_mlp_tc_kernel|node_proj_tc_kernel" \
  --launch-skip 0 --launch-count 12 -o gpurun_out/r02b_kernels_full python scratch/step_profile.py f16 train-mcr > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02b_kernels_full.ncu-rep; tail -3 gpurun_out/ncu_full.log
