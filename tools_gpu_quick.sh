#!/bin/bash
# quick GPU check: tensor-core block first (bounded), then full parity, then bench
mkdir -p gpurun_out
TAG=${1:-q}
timeout 120 python -m pytest tests/test_cell_gpu.py -x -q -k "tf32x3" 2>&1 | tail -15 | tee gpurun_out/pytest_tc_$TAG.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
