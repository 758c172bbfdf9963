#!/bin/bash
# CUDA-graph training step: equality test, then eager vs graph timing (mc and mcr, tc contractions)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "cuda_graph_equals_eager or train_step_rmsprop or edge_mlp_backward" > gpurun_out/train_graph_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/train_graph_tests.log
tail -25 gpurun_out/train_graph_tests.log
for v in mc mcr; do for gr in "" "--train-graph"; do
  sc=1024; [ $v = mcr ] && sc=512
  timeout 300 python bench.py --mode train --variant $v --train-gemm tc $gr --scenes $sc --steps 5 --warmup 3 \
    > gpurun_out/train_${v}_tc${gr:+_graph}.json 2> gpurun_out/train_${v}_tc${gr:+_graph}.err
  echo "train $v tc $gr rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/train_${v}_tc${gr:+_graph}.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches","loss_first","loss_last")})
except Exception as e:
    print("no line", e)
PY
  tail -3 gpurun_out/train_${v}_tc${gr:+_graph}.err
done; done
