#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list, ncu full capture. Outputs -> gpurun_out/
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_$TAG.log
timeout 900 python bench.py 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
tail -3 gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 460 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
tail -2 gpurun_out/ncu_list_$TAG.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_|support_" -s 125 -c 14 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --batch 1024 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/
