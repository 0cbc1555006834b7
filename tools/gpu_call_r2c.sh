#!/bin/bash
# staged backward (world size 1) + dW chain length 4: halo tests first (bounded), margins, full parity, config 3 number
mkdir -p gpurun_out
T=${1:-r2c}
timeout 240 python -m pytest tests/test_halo_gpu.py -m gpu -x -q > gpurun_out/pytest_halo_$T.log 2>&1; tail -1 gpurun_out/pytest_halo_$T.log; grep -E "^E  |FAILED" gpurun_out/pytest_halo_$T.log | head -12
timeout 200 python tools/margins.py 2 2>&1 | grep -v Warn | grep rep | tee gpurun_out/margins_$T.jsonl
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log; grep -E "^E  |FAILED" gpurun_out/pytest_$T.log | head
timeout 400 python tools/bench_configs.py config3 > gpurun_out/configs_$T.jsonl 2> gpurun_out/configs_$T.err; cat gpurun_out/configs_$T.jsonl | cut -c1-1100
