#!/bin/bash
# f16 default everywhere: tests touched, c1 / c5 / mcr default lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_realdata.py tests/test_gpu_mirrors.py -q -x -k "f16 or realdata or mirror or forecast" 2>&1 | tail -3
for c in c1 c5; do
  timeout 600 python bench.py --config $c > gpurun_out/realdata_${c}_f16.json 2> gpurun_out/realdata_${c}.err; echo "$c rc=$?"
done
timeout 600 python bench.py --variant mcr --steps 20 --parity-scenes 256 > gpurun_out/bench_mcr_f16.json 2> gpurun_out/bench_mcr_f16.err; echo "mcr rc=$?"
timeout 600 python bench.py --agents 256 --scenes 1024 --steps 20 --parity-scenes 64 > gpurun_out/bench_n256_f16.json 2> gpurun_out/bench_n256_f16.err; echo "n256 rc=$?"
python - <<'PY'
import json
for f in ("bench_mcr_f16","bench_n256_f16"):
    d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]); q=d["ade_fde"]["delta_vs_oracle"]
    print(f, round(d["value"]/1e6,2), d["dtype"][:3], d["config"]["precision_mode"], {k:q[k] for k in ("max_abs_d_ade","max_abs_d_fde","scenes_with_a_flipped_neighbour","within_1e-3","best_k_equal_frac")})
for c in ("c1","c5"):
    d=json.loads([l for l in open(f"gpurun_out/realdata_{c}_f16.json") if l.startswith("{")][-1])
    print(c, d["dtype"], {k:(round(v["ade"],5), round(v["fde"],5), v["oracle"]["max_abs_d_ade"], v["oracle"]["max_abs_d_fde"], v["oracle"]["best_k_equal_frac"]) for k,v in d["splits"].items()})
PY
