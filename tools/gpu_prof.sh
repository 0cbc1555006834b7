#!/bin/bash
# usage: tools/gpu_prof.sh TAG KERNEL_REGEX SKIP COUNT [BATCH] -- one --set full capture (with source) of selected kernels
mkdir -p gpurun_out
TAG=$1; RX=$2; SKIP=${3:-30}; CNT=${4:-2}; BATCH=${5:-4096}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --batch $BATCH --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/prof_$TAG*
