#!/bin/bash
# f16 as the benched mode: tests at the tightened tolerances, smoke, default bench line (parity on 1024 scenes), c1 / c5
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_realdata.py -q -x -k "f16 or realdata or c1 or c5" 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 900 python bench.py --parity-scenes 1024 > gpurun_out/bench_default_f16_par1024.json 2> gpurun_out/bench_default.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_default_f16_par1024.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["dtype"], d["config"]["precision_mode"])
print({k:(round(v["value"]/1e6,1), v["max_abs_d_ade_vs_oracle"], v["max_abs_d_fde_vs_oracle"], v["within_1e-3"]) for k,v in d["modes"].items()})
print(d["ade_fde"]["delta_vs_oracle"])
PY
for c in c1 c5; do
  timeout 600 python bench.py --config $c > gpurun_out/realdata_${c}_f16.json 2> gpurun_out/realdata_${c}.err
  echo "$c rc=$?"; tail -c 1500 gpurun_out/realdata_${c}_f16.json
done
