#!/bin/bash
# remaining seconds of the round: graph replay of the headline step at B = 4096, two small-batch points with the current kernels
mkdir -p gpurun_out
T=${1:-r2g}
timeout 40 python bench.py --cuda-graph --no-cpu-baseline --no-roofline 2> gpurun_out/bench_graph_$T.err | tee gpurun_out/bench_graph_$T.json | cut -c1-170
for B in 512 32; do
timeout 25 python bench.py --batch $B --no-cpu-baseline --no-roofline 2>> gpurun_out/bench_small_$T.err | tee -a gpurun_out/bench_small_$T.jsonl | cut -c1-170
done
