#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
echo "== gemm + train tests"; timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -q -m gpu -k "gemm or train or adjoint" > $O/c9_pytest.txt 2>&1; tail -6 $O/c9_pytest.txt
echo "== train bench"; for g in fp32 tf32 tc; do timeout 600 python bench.py --mode train --scenes 1024 --steps 5 --warmup 2 --train-gemm $g > $O/c9_train_$g.json 2> $O/c9_train_$g.err; python -c "
import json; d=json.loads(open('$O/c9_train_$g.json').read().strip().splitlines()[-1]); print('$g', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],2), 'ms', d['loss_first'], d['loss_last'], d['gpu_launches'])"; done
echo "== profiles"
timeout 300 python scratch/step_profile.py bf16 bf16x3 train > $O/c9_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_modes.csv python scratch/step_profile.py bf16 bf16x3 train > $O/c9_ncu1.log 2>&1
timeout 300 python scratch/step_profile.py bf16 bf16x3 train > $O/c9_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"rollout_tc_kernel|decode_score_kernel|gsk_cell_tc_kernel|gemm_tf32_kernel|aggregate_kernel" -s 2 -c 9 -o $O/r02_kernels_full python scratch/step_profile.py bf16 bf16x3 train > $O/c9_ncu2.log 2>&1
ls -la $O | grep r02; tail -3 $O/c9_ncu1.log $O/c9_ncu2.log
