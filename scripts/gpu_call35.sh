#!/bin/bash
# neighbour flips vs outliers: parity sample with the adjacency comparison, 256 and 2048 scenes
mkdir -p gpurun_out
for n in 256 2048; do
  timeout 1200 python bench.py --steps 20 --parity-scenes $n > gpurun_out/bench_f16_par$n.json 2> gpurun_out/bench_par$n.err
  echo "rc=$?"; tail -2 gpurun_out/bench_par$n.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_f16_par$n.json") if l.startswith("{")][-1])
for k,v in d["modes"].items():
    print($n, k, {kk:v.get(kk) for kk in ("max_abs_d_fde_vs_oracle","within_1e-3","scenes_with_an_agent_over_1e-3","scenes_with_a_flipped_neighbour","within_1e-3_where_adjacency_agrees","max_abs_d_fde_where_adjacency_agrees")})
print(d["ade_fde"]["delta_vs_oracle"])
PY
done
