#!/bin/bash
# dW chain length 8: config 3 parity twice in fresh processes + error margins, full parity, configs 3 / 5 numbers
mkdir -p gpurun_out
T=${1:-r2b}
for i in 1 2; do timeout 200 python -m pytest tests/test_cell_gpu.py -m gpu -x -q -k "config3" 2>&1 | tail -1; done
timeout 200 python tools/margins.py 2 2>&1 | grep -v Warn | tee gpurun_out/margins_$T.jsonl
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log; grep -E "^E  |FAILED" gpurun_out/pytest_$T.log | head
timeout 400 python tools/bench_configs.py config3 config5 > gpurun_out/configs_$T.jsonl 2> gpurun_out/configs_$T.err; cat gpurun_out/configs_$T.jsonl | cut -c1-1100
