#!/bin/bash
# compute-sanitizer memcheck over the smallest parity cases of every kernel family + the LongC-like shape
mkdir -p gpurun_out
T=${1:-r1y}
timeout 300 python -m pytest tests/test_cell_gpu.py -m gpu -x -q -k "shape10" 2>&1 | tail -2
timeout 600 python tools/bench_configs.py config5 > gpurun_out/config5_$T.jsonl 2> gpurun_out/config5_$T.err; cut -c1-1200 gpurun_out/config5_$T.jsonl
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_cell_gpu.py -m gpu -x -q \
   -k "golden and (tiny or sf_din16 or sf_din1) or (oracle_dense and (shape0 or shape4 or shape7 or shape8) and None) or (oracle_csr and shape0) or support_apply" \
   > gpurun_out/memcheck_$T.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/memcheck_$T.log | head -12
