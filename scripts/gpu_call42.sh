#!/bin/bash
# fp16 operands in the per-step kernels: tests, then mcr / N = 256 bench lines in f16
mkdir -p gpurun_out
MMT_RECORD_ERRORS=gpurun_out/f16_step_errors.jsonl timeout 900 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "f16 or forecast_bf16 or gsk_cell or rollout_bf16 or edge_mlp or invalid_slots" > gpurun_out/f16_step_tests.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/f16_step_tests.log
for v in "--variant mcr" "--agents 256 --scenes 1024"; do
  timeout 600 python bench.py $v --prec f16 --steps 20 --no-modes --parity-scenes 64 > gpurun_out/bench_f16_step.json 2> gpurun_out/bench_f16_step.err
  echo "rc=$?"; tail -2 gpurun_out/bench_f16_step.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_f16_step.json") if l.startswith("{")][-1])
    q=d["ade_fde"]["delta_vs_oracle"]
    print("$v", round(d["value"]/1e6,2), "M/s", round(d["ms_per_step"],3), d["dtype"][:3], {k:q[k] for k in ("max_abs_d_ade","max_abs_d_fde","scenes_with_a_flipped_neighbour","max_abs_d_fde_where_adjacency_agrees","best_k_equal_frac")})
except Exception as e:
    print("$v no line", e)
PY
done
