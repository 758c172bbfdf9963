#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/c12_errors.jsonl
MMT_RECORD_ERRORS=$PWD/$O/c12_errors.jsonl timeout 1500 python -m pytest tests -q -m gpu > $O/c12_pytest.txt 2>&1; tail -12 $O/c12_pytest.txt
echo "== c2"; timeout 900 python bench.py --config c2 --steps 30 --warmup 5 --train-gemm tc > $O/c12_c2.json 2> $O/c12_c2.err; echo rc=$?; tail -c 2500 $O/c12_c2.json; tail -3 $O/c12_c2.err
