#!/bin/bash
# usage: tools/gpu_profile.sh TAG [BATCH] [list|full|both] -- ncu launch list of one bench command and/or one
# --set full capture of the two gate-convolution kernels (last 8 forward + first 8 backward-dx launches of step 4).  Outputs -> gpurun_out/
mkdir -p gpurun_out
TAG=${1:-r1}; BATCH=${2:-4096}; WHAT=${3:-both}
if [ "$WHAT" != "full" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --batch $BATCH --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
tail -2 gpurun_out/ncu_list_$TAG.log
fi
if [ "$WHAT" != "list" ]; then
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_bwd_dx|tc_conv_fwd" -s 328 -c 16 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --batch $BATCH --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
SZ=$(stat -c %s gpurun_out/prof_$TAG.ncu-rep); echo "report bytes $SZ"
if [ "$SZ" -gt 45000000 ]; then rm gpurun_out/prof_$TAG.ncu-rep; echo "report too large for gpurun_out: kept the raw csv only"; fi
fi
ls -la gpurun_out/ | tail -8
