#!/bin/bash
# packed training path: tests, then the training bench lines (mc 1024 scenes, mcr 512 scenes, c2) eager and as one CUDA graph
mkdir -p gpurun_out
MMT_RECORD_ERRORS=gpurun_out/train_packed_errors.jsonl timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_realdata.py -q -x -k "train or glue or edge_mlp_backward" > gpurun_out/train_packed_tests.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/train_packed_tests.log
run() { # name, args
  timeout 300 python bench.py --mode train --train-gemm tc $2 --steps 5 --warmup 3 > gpurun_out/$1.json 2> gpurun_out/$1.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$1.json") if l.startswith("{")][-1])
    print("$1", round(d["ms_per_step"],2), "ms", round(d["value"]/1e6,3), "M/s launches", d["gpu_launches"], d["loss_first"], d["loss_last"])
except Exception as e:
    print("$1 no line", e); print(open("gpurun_out/$1.err").read()[-1500:])
PY
}
run train_mc_tc_packed "--variant mc --scenes 1024"
run train_mc_tc_packed_graph "--variant mc --scenes 1024 --train-graph"
run train_mcr_tc_packed "--variant mcr --scenes 512"
run train_mcr_tc_packed_graph "--variant mcr --scenes 512 --train-graph"
timeout 600 python bench.py --config c2 --train-gemm tc --train-graph --steps 10 --warmup 3 > gpurun_out/c2_tc_packed_graph.json 2> gpurun_out/c2_tc_packed_graph.err; echo "c2 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/c2_tc_packed_graph.json") if l.startswith("{")][-1])
print("c2", round(d["ms_per_step"],2), "ms per pair", {k:round(v["ms_per_step"],2) for k,v in d["per_table"].items()}, d["loss_first"], d["loss_last"])
h=d["held_out_zara01_best_of_20"]; print({k:(round(h[k]["ade"],4), round(h[k]["fde"],4)) for k in ("before","after")})
PY
