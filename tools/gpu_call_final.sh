#!/bin/bash
# final single-GPU session of a round: parity, headline bench + cpu baseline, reference arm, ncu list + full capture, traces
mkdir -p gpurun_out
T=${1:-r1q}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py 2> gpurun_out/bench_$T.err > gpurun_out/bench_$T.json; tail -2 gpurun_out/bench_$T.err; cut -c1-300 gpurun_out/bench_$T.json
timeout 400 python bench.py --impl reference 2> gpurun_out/ref_$T.err > gpurun_out/ref_$T.json; cut -c1-200 gpurun_out/ref_$T.json
bash tools/gpu_profile.sh $T 4096 both
timeout 200 python tools/trace_conv.py 2048 16 > gpurun_out/trace_$T.txt 2>&1
STC_OPT=9 timeout 200 python tools/trace_conv.py 2048 16 > gpurun_out/trace_${T}_smemA.txt 2>&1
timeout 300 python tools/bench_configs.py support > gpurun_out/support_$T.jsonl 2>/dev/null; cut -c1-200 gpurun_out/support_$T.jsonl | head -4
