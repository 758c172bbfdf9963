#!/bin/bash
# MMT_PREC_F16: fused rollout with fp16 operands -- parity tests, then the bench line with every mode beside it
mkdir -p gpurun_out
MMT_RECORD_ERRORS=gpurun_out/f16_errors.jsonl timeout 600 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "f16 or forecast_bf16_tensor_core or rollout_bf16" > gpurun_out/f16_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f16_tests.log
tail -15 gpurun_out/f16_tests.log
timeout 600 python bench.py --prec f16 --steps 50 --warmup 5 > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err
echo "bench f16 rc=$?"; tail -3 gpurun_out/bench_f16.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_f16.json") if l.startswith("{")][-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["dtype"])
    print(json.dumps(d["modes"], indent=0))
    print(json.dumps(d["roofline"], indent=0)[:600])
    print(d["ade_fde"]["delta_vs_oracle"])
except Exception as e:
    print("no line", e)
PY
