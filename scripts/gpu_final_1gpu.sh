#!/bin/bash
# what the driver runs at round end on one GPU: GPU test suite, smoke(), the default bench line and the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -q -m gpu > $O/final_pytest.txt 2>&1; tail -3 $O/final_pytest.txt
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | grep "^smoke"
echo "== reference arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/final_ref.json 2>/dev/null; head -c 300 $O/final_ref.json; echo
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/final_bench.json 2> $O/final_bench.err; echo rc=$?; python -c "
import json; d=json.loads(open('$O/final_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], 'pairwise', d['roofline_pairwise']['frac'], 'decode ms', d['roofline_decode']['ms_per_launch'])
print(json.dumps(d['modes'])); print(d['cpu_baseline'])"
