#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "cuda_graph_equals_eager" 2>&1 | tail -2
timeout 600 python bench.py --config c2 --train-gemm tc --train-graph --steps 10 --warmup 3 > gpurun_out/c2_tc_graph.json 2> gpurun_out/c2_tc_graph.err; echo "c2 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/c2_tc_graph.json") if l.startswith("{")][-1])
print(round(d["ms_per_step"],2), "ms per pair of steps", round(d["value"]/1e3,1), "k/s", {k:round(v["ms_per_step"],2) for k,v in d["per_table"].items()}, d["loss_first"], d["loss_last"])
h=d["held_out_zara01_best_of_20"]; print({k:(round(h[k]["ade"],4), round(h[k]["fde"],4)) for k in ("before","after")})
PY
