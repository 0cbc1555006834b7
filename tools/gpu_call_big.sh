#!/bin/bash
# wide-hidden-state forward kernel: bounded parity run first (a hang must not eat the budget), then the side configs
mkdir -p gpurun_out
T=${1:-r1r}
timeout 180 python -m pytest tests/test_cell_gpu.py -m gpu -x -q -k "shape4 or shape5 or shape6 or shape7" > gpurun_out/pytest_big_$T.log 2>&1 || { tail -40 gpurun_out/pytest_big_$T.log; echo BIG TESTS FAILED; exit 1; }
tail -1 gpurun_out/pytest_big_$T.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log
timeout 600 python tools/bench_configs.py config3 config4 > gpurun_out/configs_$T.jsonl 2> gpurun_out/configs_$T.err; cat gpurun_out/configs_$T.jsonl | cut -c1-900
