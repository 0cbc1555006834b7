#!/bin/bash
# last call of the round: full parity (incl. the partitioned-stack test), smoke, headline bench without the CPU leg
mkdir -p gpurun_out
T=${1:-r2f}
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log; grep -E "^E  |FAILED" gpurun_out/pytest_$T.log | head
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 120 python bench.py --no-cpu-baseline 2> gpurun_out/bench_$T.err > gpurun_out/bench_$T.json; tail -2 gpurun_out/bench_$T.err; cut -c1-400 gpurun_out/bench_$T.json
