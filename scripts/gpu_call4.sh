#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
echo "== ro_time"; timeout 900 python scratch/ro_time.py all > $O/c4_ro_time.txt 2>&1; grep "ms/launch" $O/c4_ro_time.txt
echo "== gpu tests"; timeout 1200 python -m pytest tests -q -m gpu > $O/c4_pytest.txt 2>&1; tail -25 $O/c4_pytest.txt
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-modes > $O/c4_bench.json 2> $O/c4_bench.err; echo rc=$?; python -c "
import json; d=json.loads(open('$O/c4_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline_decode']['ms_per_launch'], d['gpu_launches'])"; tail -3 $O/c4_bench.err
