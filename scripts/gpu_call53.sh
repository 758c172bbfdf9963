#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "train or glue" 2>&1 | tail -2
run() {
  timeout 300 python bench.py --mode train --train-gemm tc $2 --steps 5 --warmup 3 > gpurun_out/$1.json 2> gpurun_out/$1.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/$1.json") if l.startswith("{")][-1])
print("$1", round(d["ms_per_step"],2), "ms", round(d["value"]/1e6,3), "M/s", d["loss_first"], d["loss_last"])
PY
}
run train_mc_tc_packed "--variant mc --scenes 1024"
run train_mc_tc_packed_graph "--variant mc --scenes 1024 --train-graph"
run train_mcr_tc_packed "--variant mcr --scenes 512"
run train_mcr_tc_packed_graph "--variant mcr --scenes 512 --train-graph"
timeout 300 python scratch/train_kernel_count.py mc > gpurun_out/train_kernels_mc.txt 2>&1
timeout 300 python scratch/train_kernel_count.py mcr > gpurun_out/train_kernels_mcr.txt 2>&1
grep "kernels per step" gpurun_out/train_kernels_mc.txt gpurun_out/train_kernels_mcr.txt
