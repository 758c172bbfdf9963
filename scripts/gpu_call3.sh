#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
echo "== gpu tests"; timeout 1200 python -m pytest tests -q -m gpu > $O/c3_pytest.txt 2>&1; tail -25 $O/c3_pytest.txt
echo "== ro_time"; timeout 600 python scratch/ro_time.py all > $O/c3_ro_time.txt 2>&1; cat $O/c3_ro_time.txt
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > $O/c3_bench.json 2> $O/c3_bench.err; echo rc=$?; python -c "
import json; d=json.loads(open('$O/c3_bench.json').read().strip().splitlines()[-1]); print(json.dumps(d['modes'],indent=1)); print(d['value'], d['e2e'], d['roofline']['ms_per_launch'])"; tail -3 $O/c3_bench.err
