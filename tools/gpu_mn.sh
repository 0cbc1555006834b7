#!/bin/bash
# MN-major tcgen05 operand layouts: which variant of the self-test kernel reproduces fp64?
mkdir -p gpurun_out
for v in 0 1 2; do
  echo "== MN variant $v" | tee -a gpurun_out/mn_variants.log
  STC_TC_TEST_MODE=2 STC_TC_MN_VARIANT=$v timeout 120 python -m pytest tests/test_cell_gpu.py -q -k "tf32x3" 2>&1 | grep -E "passed|failed|Error" | tee -a gpurun_out/mn_variants.log
done
timeout 300 python tools/tc_error.py 2>&1 | tee gpurun_out/tc_error.log
