#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | grep -v "^\[multi" | tail -5
echo "== tests"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
echo "== train"; for g in tf32 tc; do timeout 600 python bench.py --mode train --scenes 1024 --steps 5 --warmup 2 --train-gemm $g > $O/c13_train_$g.json 2> $O/c13_train_$g.err; python -c "
import json; d=json.loads(open('$O/c13_train_$g.json').read().strip().splitlines()[-1]); print('$g', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],2), 'ms', d['loss_first'], d['loss_last'], d['gpu_launches'])"; done
echo "== C4"; timeout 600 python bench.py --scenes 1024 --agents 256 --steps 10 --warmup 3 --no-modes --parity-scenes 32 > $O/c13_c4.json 2> $O/c13_c4.err; python -c "
import json; d=json.loads(open('$O/c13_c4.json').read().strip().splitlines()[-1]); print('C4', round(d['value']/1e6,2), 'M/s', round(d['ms_per_step'],3), d['ade_fde']['delta_vs_oracle'])"; tail -2 $O/c13_c4.err
echo "== mcr"; timeout 600 python bench.py --variant mcr --steps 10 --warmup 3 --no-modes --parity-scenes 64 > $O/c13_mcr.json 2> $O/c13_mcr.err; python -c "
import json; d=json.loads(open('$O/c13_mcr.json').read().strip().splitlines()[-1]); print('mcr', round(d['value']/1e6,2), 'M/s', round(d['ms_per_step'],3), d['ade_fde']['delta_vs_oracle'])"; tail -2 $O/c13_mcr.err
