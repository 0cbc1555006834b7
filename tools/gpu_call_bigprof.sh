#!/bin/bash
# ncu --set full capture of the wide-hidden-state forward kernel at the config 3 shape (N = 4096, C = 16, F = 64)
mkdir -p gpurun_out
T=${1:-r1r}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_fwd_big" -s 4 -c 2 \
    -o gpurun_out/prof_big_$T -f python tools/bench_configs.py config3 > gpurun_out/ncu_big_$T.log 2>&1
tail -3 gpurun_out/ncu_big_$T.log
ncu -i gpurun_out/prof_big_$T.ncu-rep --page raw --csv > gpurun_out/prof_big_${T}_raw.csv 2>/dev/null
ls -la gpurun_out/prof_big_$T*
