#!/bin/bash
# round 2, GPU call 1 (2 GPUs): root-cause experiment for the multi-GPU CUDA 719, then the driver's own launch line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/c1_gpus.txt
echo "== ro_race" ; timeout 600 python scratch/ro_race.py run > $O/c1_ro_race.txt 2>&1; cat $O/c1_ro_race.txt | tail -8
run2() {  # $1 tag, rest: env assignments
  tag=$1; shift
  env "$@" NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$O/c1_nccl_${tag}.%p.log timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
     --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c1_bench2_${tag}.out 2> $O/c1_bench2_${tag}.err
  echo "bench2 $tag rc=$?"; tail -c 600 $O/c1_bench2_${tag}.out; grep -h "FAILED in stage\|mmt_last" $O/c1_bench2_${tag}.err | head -4
  for f in $O/rank*.err; do [ -f "$f" ] && mv "$f" "$O/c1_${tag}_$(basename $f)"; done
}
for i in 1 2 3; do run2 fix$i A=1; done
for i in 1 2 3; do run2 pre$i MMT_LIB=$PWD/build/libmmt_norder.so; done
rm -f $O/c1_nccl_*.log.keep; ls $O | head -50
echo "== gpu tests"; timeout 900 python -m pytest tests -x -q -m gpu > $O/c1_pytest.txt 2>&1; tail -5 $O/c1_pytest.txt
