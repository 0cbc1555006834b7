#!/bin/bash
# measurement session: headline bench (+cpu baseline), reference arm, ncu launch list + full capture, side configs, sweep
mkdir -p gpurun_out
T=${1:-r1m}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader | tee gpurun_out/gpu_$T.txt
timeout 600 python bench.py 2> gpurun_out/bench_$T.err > gpurun_out/bench_$T.json; tail -2 gpurun_out/bench_$T.err; cut -c1-400 gpurun_out/bench_$T.json
timeout 400 python bench.py --impl reference 2> gpurun_out/ref_$T.err > gpurun_out/ref_$T.json; cut -c1-200 gpurun_out/ref_$T.json
bash tools/gpu_profile.sh $T 4096 both
timeout 600 python tools/bench_configs.py support config3 config4 > gpurun_out/configs_$T.jsonl 2> gpurun_out/configs_$T.err; cut -c1-230 gpurun_out/configs_$T.jsonl
timeout 900 python tools/bench_configs.py sweep > gpurun_out/sweep_$T.jsonl 2> gpurun_out/sweep_$T.err; cut -c1-200 gpurun_out/sweep_$T.jsonl
