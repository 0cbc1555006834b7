#!/bin/bash
# quick GPU check: tensor-core blocks first (bounded), then full parity, then bench
mkdir -p gpurun_out
TAG=${1:-q}
timeout 180 python -m pytest tests/test_cell_gpu.py -x -q -k "tf32x3" 2>&1 | tail -3 | tee gpurun_out/pytest_tc_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/pytest_$TAG.log; grep -E "passed|failed|Error" gpurun_out/pytest_$TAG.log | head
timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench_$TAG.err > gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1))
for k,v in d['kernel_breakdown'].items(): print(' ',k,{a:round(b,3) for a,b in v.items()})
PY
