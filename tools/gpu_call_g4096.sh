#!/bin/bash
# side workload of BASELINE configs[2] (N=4096, C=16, F=64, T=12) + a sanity pass of the headline line and the tests
mkdir -p gpurun_out
T=${1:-r1w}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log
timeout 600 python bench.py --workload g4096 --steps 5 --warmup 3 2> gpurun_out/bench_g4096_$T.err > gpurun_out/bench_g4096_$T.json; tail -3 gpurun_out/bench_g4096_$T.err; cut -c1-1500 gpurun_out/bench_g4096_$T.json
timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench_$T.err > gpurun_out/bench_$T.json; cut -c1-200 gpurun_out/bench_$T.json
