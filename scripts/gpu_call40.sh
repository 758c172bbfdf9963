#!/bin/bash
# g2k_lstm_mcr inference (bf16 per-step tensor-core kernels): bench line and launch list at C3; same for N = 256 (mc)
mkdir -p gpurun_out
timeout 600 python bench.py --variant mcr --prec bf16 --steps 20 --no-modes --parity-scenes 64 > gpurun_out/bench_mcr.json 2> gpurun_out/bench_mcr.err; echo "rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_mcr.json') if l.startswith('{')][-1]); print('mcr', d['value']/1e6, d['ms_per_step'], d['gpu_launches'], d['ade_fde']['delta_vs_oracle']['max_abs_d_ade'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mcr_launches.csv \
  python bench.py --variant mcr --prec bf16 --steps 2 --warmup 3 --no-modes --no-graph --parity-scenes 8 > gpurun_out/mcr_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --agents 256 --scenes 1024 --prec bf16 --steps 20 --no-modes --parity-scenes 16 > gpurun_out/bench_n256.json 2> gpurun_out/bench_n256.err; echo "rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n256.json') if l.startswith('{')][-1]); print('n256', d['value']/1e6, d['ms_per_step'], d['gpu_launches'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n256_launches.csv \
  python bench.py --agents 256 --scenes 1024 --prec bf16 --steps 2 --warmup 3 --no-modes --no-graph --parity-scenes 8 > gpurun_out/n256_ncu.log 2>&1; echo "ncu rc=$?"
