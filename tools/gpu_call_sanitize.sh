#!/bin/bash
# compute-sanitizer memcheck over the smallest parity cases of every kernel family + the LongC-like shape
mkdir -p gpurun_out
T=${1:-r1y}
timeout 300 python -m pytest "tests/test_cell_gpu.py::test_cell_matches_oracle_dense[None-shape8]" "tests/test_cell_gpu.py::test_cell_matches_oracle_dense[relu-shape8]" -x -q 2>&1 | tail -2
timeout 600 python tools/bench_configs.py config5 > gpurun_out/config5_$T.jsonl 2> gpurun_out/config5_$T.err; cut -c1-1200 gpurun_out/config5_$T.jsonl
P=tests/test_cell_gpu.py
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -x -q \
   "$P::test_cell_matches_reference_golden[tiny]" "$P::test_cell_matches_reference_golden[sf_din16]" "$P::test_cell_matches_reference_golden[sf_din1]" \
   "$P::test_cell_matches_oracle_dense[None-shape1]" "$P::test_cell_matches_oracle_dense[None-shape4]" "$P::test_cell_matches_oracle_dense[None-shape6]" \
   "$P::test_cell_matches_oracle_dense[None-shape7]" "$P::test_cell_matches_oracle_dense[None-shape8]" "$P::test_cell_matches_oracle_csr[shape0]" \
   "$P::test_support_apply_matches_oracle[dense]" "$P::test_support_apply_matches_oracle[csr]" \
   > gpurun_out/memcheck_$T.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|rror:" gpurun_out/memcheck_$T.log | head -12
