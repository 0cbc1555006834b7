#!/bin/bash
# quick iteration: parity, bench A/B over STC_OPT values, phase traces
mkdir -p gpurun_out
T=${1:-r1g}
OPTS=${2:-"1 5"}
timeout 150 python -m pytest tests/test_cell_gpu.py -m gpu -x -q -k "support_apply or sf_din16 or tiny" > gpurun_out/pytest_quick_$T.log 2>&1 || { tail -30 gpurun_out/pytest_quick_$T.log; echo QUICK TESTS FAILED; exit 1; }
tail -1 gpurun_out/pytest_quick_$T.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -3 gpurun_out/pytest_$T.log
for OPT in $OPTS; do
  STC_OPT=$OPT timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench_${T}_opt$OPT.err > gpurun_out/bench_${T}_opt$OPT.json
  tail -2 gpurun_out/bench_${T}_opt$OPT.err
done
python - <<PY
import json
for opt in "$OPTS".split():
    try:
        d = json.load(open(f"gpurun_out/bench_${T}_opt{opt}.json"))
        kb = d["kernel_breakdown"]
        print("opt", opt, round(d["value"]), "samples/s", round(d["ms_per_step"], 2), "ms | e2e", round(d["e2e"]["value"]), "|",
              " ".join(f"{k}={v['ms_per_step']:.1f}" for k, v in list(kb.items())[:6]))
    except Exception as e:
        print(opt, "failed", e)
PY
timeout 200 python tools/trace_conv.py 2048 16 > gpurun_out/trace_${T}.txt 2>&1
timeout 200 python tools/trace_conv.py 2048 1 > gpurun_out/trace_${T}_din1.txt 2>&1
cat gpurun_out/trace_${T}.txt
