#!/bin/bash
# round 2, GPU call 2 (1 GPU): test suite, default bench line (modes + parity deltas), real-data configs, trap-record self-test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
echo "== gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -x > $O/c2_pytest.txt 2>&1; tail -15 $O/c2_pytest.txt
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > $O/c2_bench.json 2> $O/c2_bench.err; echo rc=$?; tail -c 3000 $O/c2_bench.json; tail -5 $O/c2_bench.err
echo "== c1"; timeout 600 python bench.py --config c1 > $O/c2_c1.json 2> $O/c2_c1.err; echo rc=$?; cat $O/c2_c1.json; tail -3 $O/c2_c1.err
echo "== c5"; timeout 900 python bench.py --config c5 > $O/c2_c5.json 2> $O/c2_c5.err; echo rc=$?; cat $O/c2_c5.json; tail -3 $O/c2_c5.err
echo "== trap self-test"; MMT_RO_FLAGS=1024 timeout 300 python scratch/ro_race.py child 2>&1 | tail -2 > $O/c2_trap_selftest.txt; cat $O/c2_trap_selftest.txt
