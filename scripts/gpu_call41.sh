#!/bin/bash
# wide configuration of the blocked bf16 cell kernel: equality with the narrow one, then mcr / N = 256 bench lines A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "wide_equals_narrow or forecast_bf16 or gsk_cell" 2>&1 | tail -4
for w in -1 1; do
  MMT_CELL_WIDE=$w timeout 600 python bench.py --variant mcr --prec bf16 --steps 20 --no-modes --parity-scenes 64 > gpurun_out/bench_mcr_w$w.json 2> gpurun_out/bench_mcr_w$w.err
  MMT_CELL_WIDE=$w timeout 600 python bench.py --agents 256 --scenes 1024 --prec bf16 --steps 20 --no-modes --parity-scenes 16 > gpurun_out/bench_n256_w$w.json 2> gpurun_out/bench_n256_w$w.err
  python - <<PY
import json
for f in ("mcr","n256"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench_{f}_w$w.json") if l.startswith("{")][-1])
        print("wide=$w", f, round(d["value"]/1e6,2), "M/s", round(d["ms_per_step"],3), "ms", "cell kernel ms", d["roofline"]["ms_per_launch"], d["ade_fde"]["delta_vs_oracle"]["max_abs_d_ade"])
    except Exception as e:
        print("wide=$w", f, "no line", e)
PY
done
