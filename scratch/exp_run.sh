#!/bin/bash
# experiment driver: microbenchmark + kernel time of the library variants given as arguments
mkdir -p gpurun_out
./scratch/lds_bench > gpurun_out/lds_bench.txt 2>&1
for v in "$@"; do
  if [ "$v" = base ]; then unset MMT_LIB; else export MMT_LIB=$PWD/multimodaltraj_2_b200/lib/libmmt_$v.so; fi
  echo "== $v" >> gpurun_out/exp_times.txt
  timeout 300 python scratch/ro_timeline.py 2>&1 | tail -1 >> gpurun_out/exp_times.txt
done
unset MMT_LIB
cat gpurun_out/lds_bench.txt gpurun_out/exp_times.txt
