#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"edge_mlp_bwd_tc_kernel|attention_score_grad_kernel" \
  --launch-skip 20 --launch-count 4 -o gpurun_out/r02b_edge_bwd_full python scratch/step_profile.py train-mcr > gpurun_out/ncu_full2.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02b_edge_bwd_full.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_mcr_launches.csv \
  python bench.py --mode train --variant mcr --train-gemm tc --scenes 512 --steps 1 --warmup 3 > gpurun_out/train_mcr_ncu.log 2>&1
echo "ncu2 rc=$?"; wc -l gpurun_out/train_mcr_launches.csv
