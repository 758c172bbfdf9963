#!/bin/bash
# where the training step's GPU time goes: ncu launch list of `bench.py --mode train --train-gemm tc` (mc, 1024 scenes)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "cuda_graph_equals_eager" 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
  python bench.py --mode train --variant mc --train-gemm tc --scenes 1024 --steps 1 --warmup 3 > gpurun_out/train_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/train_launches.csv
