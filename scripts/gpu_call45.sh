#!/bin/bash
# BASELINE configs[1] (g2k_lstm_mcr training on the real zara1 leave-one-out tables) on the edge-backward kernels
mkdir -p gpurun_out
for g in tc fp32; do
  timeout 600 python bench.py --config c2 --train-gemm $g --steps 10 --warmup 3 > gpurun_out/c2_$g.json 2> gpurun_out/c2_$g.err; echo "c2 $g rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/c2_$g.json") if l.startswith("{")][-1])
print("$g", round(d["ms_per_step"],2), "ms per pair of steps", round(d["value"]/1e3,1), "k/s", d["per_table"], d["loss_first"], d["loss_last"])
h=d["held_out_zara01_best_of_20"]; print({k:(round(h[k]["ade"],4), round(h[k]["fde"],4)) for k in ("before","after")})
PY
done
