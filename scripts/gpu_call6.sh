#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x > $O/c6_gemm.txt 2>&1; tail -30 $O/c6_gemm.txt
echo "== all gpu tests"; timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_gemm.py > $O/c6_pytest.txt 2>&1; tail -8 $O/c6_pytest.txt
echo "== mufu3"; timeout 120 ./build/mufu_bench3 > $O/c6_mufu3.txt 2>&1; cat $O/c6_mufu3.txt
echo "== ro_time"; timeout 300 python scratch/ro_time.py all 2>&1 | grep "ms/launch" | tee $O/c6_ro_time.txt
