#!/bin/bash
# error distribution of the operand modes on 256 / 2048 scenes of the benched batch
mkdir -p gpurun_out
for n in 256 2048; do
  timeout 1200 python bench.py --steps 20 --parity-scenes $n > gpurun_out/bench_f16_par$n.json 2> gpurun_out/bench_par$n.err
  echo "rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_f16_par$n.json") if l.startswith("{")][-1])
for k,v in d["modes"].items():
    print($n, k, {kk:v[kk] for kk in ("max_abs_d_ade_vs_oracle","max_abs_d_fde_vs_oracle","within_1e-3","frac_samples_within_1e-3","p9999_abs_d_fde_vs_oracle","scenes_with_an_agent_over_1e-3")})
print(d["ade_fde"]["delta_vs_oracle"])
PY
done
