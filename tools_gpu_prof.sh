#!/bin/bash
# usage: tools_gpu_prof.sh TAG KERNEL_REGEX SKIP COUNT [BATCH]
mkdir -p gpurun_out
TAG=$1; RX=$2; SKIP=${3:-30}; CNT=${4:-2}; BATCH=${5:-1024}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --batch $BATCH --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
