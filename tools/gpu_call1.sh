#!/bin/bash
# GPU session r1f: parity, headline bench, reference arm, tuning-switch A/B, phase traces, side configs, batch sweep
mkdir -p gpurun_out
T=r1f
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader | tee gpurun_out/gpu_$T.txt
ls MEASURED_PEAKS.json 2>/dev/null && cp MEASURED_PEAKS.json gpurun_out/peaks_$T.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -3 gpurun_out/pytest_$T.log
timeout 600 python bench.py 2> gpurun_out/bench_$T.err > gpurun_out/bench_$T.json; tail -2 gpurun_out/bench_$T.err
for OPT in 0 1 2 4; do
  STC_OPT=$OPT timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench_${T}_opt$OPT.err > gpurun_out/bench_${T}_opt$OPT.json
done
python - <<PY
import json
for tag in ["", "_opt0", "_opt1", "_opt2", "_opt4"]:
    try:
        d = json.load(open(f"gpurun_out/bench_$T{tag}.json"))
        kb = d["kernel_breakdown"]
        print(tag or "_all", round(d["value"]), "samples/s", round(d["ms_per_step"], 2), "ms |",
              " ".join(f"{k}={v['ms_per_step']:.1f}" for k, v in list(kb.items())[:6]))
    except Exception as e:
        print(tag, "failed", e)
PY
STC_OPT=0 timeout 200 python tools/trace_conv.py 2048 16 > gpurun_out/trace_${T}_opt0.txt 2>&1
timeout 200 python tools/trace_conv.py 2048 16 > gpurun_out/trace_${T}_all.txt 2>&1
timeout 200 python tools/trace_conv.py 2048 1 > gpurun_out/trace_${T}_all_din1.txt 2>&1
cat gpurun_out/trace_${T}_opt0.txt gpurun_out/trace_${T}_all.txt | head -90
timeout 400 python bench.py --impl reference 2> gpurun_out/ref_$T.err > gpurun_out/ref_$T.json; cat gpurun_out/ref_$T.json | cut -c1-300
timeout 600 python tools/bench_configs.py support config3 config4 > gpurun_out/configs_$T.jsonl 2> gpurun_out/configs_$T.err; tail -3 gpurun_out/configs_$T.err
cut -c1-260 gpurun_out/configs_$T.jsonl
timeout 900 python tools/bench_configs.py sweep > gpurun_out/sweep_$T.jsonl 2> gpurun_out/sweep_$T.err
cut -c1-260 gpurun_out/sweep_$T.jsonl
