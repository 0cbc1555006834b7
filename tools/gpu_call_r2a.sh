#!/bin/bash
# wide dx change check: bounded wide-shape parity first, full parity, then configs 3 / 5 cell-step numbers
mkdir -p gpurun_out
T=${1:-r2a}
timeout 240 python -m pytest tests/test_cell_gpu.py -m gpu -x -q -k "shape4 or shape5 or shape6 or shape7 or shape8 or config3" > gpurun_out/pytest_big_$T.log 2>&1 || { tail -40 gpurun_out/pytest_big_$T.log; echo BIG TESTS FAILED; exit 1; }
tail -1 gpurun_out/pytest_big_$T.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log
timeout 400 python tools/bench_configs.py config3 config5 > gpurun_out/configs_$T.jsonl 2> gpurun_out/configs_$T.err; cat gpurun_out/configs_$T.jsonl | cut -c1-1100
