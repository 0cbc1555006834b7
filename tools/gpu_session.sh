#!/bin/bash
# One GPU session = a list of stages run on the box by gpurun; everything lands in gpurun_out/ tagged with TAG.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r3a tests smoke bench refarm launches ncu'
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_session.sh r3b tests2 bench:2 halo:2'
# Stages (N = number of GPUs for the :N forms):
#   tests            pytest -m gpu (whole suite)           tests2     the multi-rank NCCL tests only
#   tests:EXPR       pytest -m gpu -k EXPR                 smoke      __graft_entry__.smoke()
#   bench[:N]        headline bench line (torchrun for N>1) refarm    bench.py --impl reference
#   bench_b:B        headline bench at batch B (no baselines)   bench_g:B  the same step replayed from a CUDA graph
#   g4096[:N[:B]]    bench.py --workload g4096             sweep      config 2 batch sweep (tools/bench_configs.py)
#   launches[:B]     ncu launch list of one bench step     ncu[:REGEX[:SKIP[:COUNT]]]  one ncu --set full capture
#   support|config3|config4|config5   tools/bench_configs.py side workloads
#   full:N:CFG:B[:stock|adam]  reference STCGNN + installed cell, DP over N GPUs (CFG = sf | longc)
#   mainpy[:EPOCHS]  the reference's Main.py for EPOCHS epochs: stock cell | installed | installed + loop hygiene
#   halo:N[:train]   tools/bench_halo.py on N GPUs         trace      clock64 phase trace of the gate convolutions
#   memcheck         compute-sanitizer over the smallest parity case of every kernel family (memcheck2: the dense N > 128 kernels)
#   dense            tools/bench_dense_support.py: dense support with N > 128 (tcgen05 vs FFMA vs cuBLAS fp32)
#   probe            what the box has (GPU, host cores / memory, reference probe)
mkdir -p gpurun_out
T=${1:?tag}; shift
RUN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port 29521 "${@:2}"; }
for stage in "$@"; do
  IFS=: read -r S A1 A2 A3 <<< "$stage"
  echo "=== stage $stage"
  case $S in
    probe)
      { nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader; nproc; free -g | head -2
        for d in "${STC_REF_DIR:-/nonexistent}" /root/reference/framework baseline/_ref/framework; do
          [ -f "$d/STC_GNN.py" ] && echo "reference found: $d" || echo "no reference at $d"; done; } | tee gpurun_out/probe_$T.txt ;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -q ${A1:+-k "$A1"} -rs > gpurun_out/pytest_$T.log 2>&1
      tail -4 gpurun_out/pytest_$T.log; grep -E "^E  |FAILED|^ERROR" gpurun_out/pytest_$T.log | head -20 ;;
    tests2)
      timeout 900 python -m pytest tests -m gpu -q -k "two_ranks or multi_rank" -rs > gpurun_out/pytest2_$T.log 2>&1
      tail -4 gpurun_out/pytest2_$T.log; grep -E "^E  |FAILED|^ERROR" gpurun_out/pytest2_$T.log | head -20 ;;
    dropin)
      timeout 900 python -m pytest tests/test_dropin_gpu.py -m gpu -q -s -rs > gpurun_out/dropin_$T.log 2>&1
      tail -4 gpurun_out/dropin_$T.log; grep -E "^E  |FAILED|^ERROR" gpurun_out/dropin_$T.log | head -20 ;;
    smoke) timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 ;;
    dense) timeout 240 python tools/bench_dense_support.py 2> gpurun_out/dense_$T.err | tee gpurun_out/dense_$T.jsonl; tail -3 gpurun_out/dense_$T.err ;;
    bench)
      N=${A1:-1}; O=gpurun_out/bench_${N}gpu_$T
      if [ "$N" = 1 ]; then timeout 900 python bench.py 2> $O.err > $O.json
      else timeout 900 bash -c "$(declare -f RUN); RUN $N bench.py --gpus $N --steps 10 --warmup 3" 2> $O.err > $O.json; fi
      tail -2 $O.err; cut -c1-400 $O.json ;;
    bench_b)
      O=gpurun_out/bench_b${A1}_$T
      timeout 600 python bench.py --batch "$A1" --no-cpu-baseline --no-eager-baseline 2> $O.err > $O.json; tail -2 $O.err; cut -c1-300 $O.json ;;
    bench_g)   # CUDA-graph replay of the whole step at batch A1
      O=gpurun_out/bench_graph_b${A1}_$T
      timeout 600 python bench.py --batch "$A1" --cuda-graph --no-cpu-baseline --no-eager-baseline --no-roofline 2> $O.err > $O.json; tail -2 $O.err; cut -c1-300 $O.json ;;
    refarm) timeout 600 python bench.py --impl reference 2> gpurun_out/ref_$T.err > gpurun_out/ref_$T.json; cut -c1-300 gpurun_out/ref_$T.json ;;
    g4096)
      N=${A1:-1}; B=${A2:-4}; O=gpurun_out/g4096_${N}gpu_b${B}_$T
      if [ "$N" = 1 ]; then timeout 900 python bench.py --workload g4096 --batch $B --steps 5 2> $O.err > $O.json
      else timeout 900 bash -c "$(declare -f RUN); RUN $N bench.py --workload g4096 --gpus $N --batch $B --steps 5 --warmup 3" 2> $O.err > $O.json; fi
      tail -2 $O.err; cut -c1-300 $O.json ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$T.csv \
        python bench.py --steps 1 --warmup 3 --batch ${A1:-4096} --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_list_$T.log 2>&1
      tail -2 gpurun_out/ncu_list_$T.log ;;
    ncu)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${A1:-tc_conv_bwd_dx|tc_conv_fwd}" -s ${A2:-328} -c ${A3:-16} \
        -o gpurun_out/prof_$T -f python bench.py --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_full_$T.log 2>&1
      tail -2 gpurun_out/ncu_full_$T.log
      ncu -i gpurun_out/prof_$T.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv 2>/dev/null
      SZ=$(stat -c %s gpurun_out/prof_$T.ncu-rep 2>/dev/null || echo 0); echo "report bytes $SZ"
      if [ "$SZ" -gt 40000000 ]; then rm gpurun_out/prof_$T.ncu-rep; echo "report too large for gpurun_out: kept the raw csv only"; fi ;;
    ncucfg)   # ncucfg:CONFIG:REGEX:SKIP:COUNT -- one ncu --set full capture of a side workload of tools/bench_configs.py
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${A2:-big}" -s ${A3:-20} -c 12 \
        -o gpurun_out/prof_${A1}_$T -f python tools/bench_configs.py $A1 > gpurun_out/ncu_${A1}_$T.log 2>&1
      tail -2 gpurun_out/ncu_${A1}_$T.log
      ncu -i gpurun_out/prof_${A1}_$T.ncu-rep --page raw --csv > gpurun_out/prof_${A1}_${T}_raw.csv 2>/dev/null
      SZ=$(stat -c %s gpurun_out/prof_${A1}_$T.ncu-rep 2>/dev/null || echo 0); echo "report bytes $SZ"
      if [ "$SZ" -gt 40000000 ]; then rm gpurun_out/prof_${A1}_$T.ncu-rep; echo "kept the raw csv only"; fi ;;
    support|config3|config4|config5|sweep)
      timeout 900 python tools/bench_configs.py $S > gpurun_out/${S}_$T.jsonl 2> gpurun_out/${S}_$T.err; cut -c1-260 gpurun_out/${S}_$T.jsonl; tail -2 gpurun_out/${S}_$T.err ;;
    halo)
      N=${A1:-2}
      IFS=: read -r _ _ _ HB <<< "$stage"
      timeout 900 bash -c "$(declare -f RUN); RUN $N tools/bench_halo.py ${HB:-2} ${A2:+--$A2}" 2>> gpurun_out/halo_${N}gpu_$T.err | tee -a gpurun_out/halo_${N}gpu_$T.jsonl | cut -c1-400
      tail -3 gpurun_out/halo_${N}gpu_$T.err ;;
    full)   # full:N:CONFIG:BATCH[:stock]  -- the unmodified reference STCGNN with the cell installed (tools/bench_full_model.py)
      N=${A1:-1}; CFG=${A2:-sf}; IFS=: read -r _ _ _ B EXTRA <<< "$stage"; B=${B:-32}
      FLAGS="--config $CFG --batch $B"; [ "$EXTRA" = stock ] && FLAGS="$FLAGS --stock"; [ "$EXTRA" = adam ] && FLAGS="$FLAGS --adam"
      if [ "$N" = 1 ]; then timeout 900 python tools/bench_full_model.py $FLAGS 2>> gpurun_out/full_$T.err | tee -a gpurun_out/full_$T.jsonl
      else timeout 900 bash -c "$(declare -f RUN); RUN $N tools/bench_full_model.py $FLAGS" 2>> gpurun_out/full_$T.err | tee -a gpurun_out/full_$T.jsonl; fi
      tail -2 gpurun_out/full_$T.err ;;
    mainpy) timeout 1200 python tools/bench_main.py --epochs ${A1:-2} 2> gpurun_out/mainpy_$T.err | tee gpurun_out/mainpy_$T.jsonl | cut -c1-400 ;;
    trace) timeout 200 python tools/trace_conv.py 2048 16 > gpurun_out/trace_$T.txt 2>&1; tail -5 gpurun_out/trace_$T.txt ;;
    memcheck)
      P=tests/test_cell_gpu.py
      timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -x -q \
        "$P::test_cell_matches_reference_golden[tiny]" "$P::test_cell_matches_reference_golden[sf_din16]" "$P::test_cell_matches_reference_golden[sf_din1]" \
        "$P::test_cell_matches_oracle_dense[None-shape1]" "$P::test_cell_matches_oracle_dense[None-shape4]" "$P::test_cell_matches_oracle_dense[None-shape6]" \
        "$P::test_cell_matches_oracle_dense[None-shape8]" "$P::test_cell_matches_oracle_csr[shape0]" \
        "$P::test_support_apply_matches_oracle[dense]" "$P::test_support_apply_matches_oracle[csr]" \
        "$P::test_inplace_and_returned_leaf_gradients_agree_and_lanes_do_not_change_results" \
        "tests/test_halo_gpu.py::test_row_subset_apply_and_halo_pack_unpack_kernels" "tests/test_halo_gpu.py::test_staged_cell_single_rank" \
        > gpurun_out/memcheck_$T.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|rror:" gpurun_out/memcheck_$T.log | head -12 ;;
    memcheck2)   # the dense N > 128 kernels (tc_support_big, block-wise tc_outer, split axpy)
      P=tests/test_cell_gpu.py
      timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -x -q \
        "$P::test_dense_support_wider_than_one_tile[130-3-shape0]" "$P::test_dense_support_wider_than_one_tile[333-1-shape3]" \
        "$P::test_cell_matches_oracle_dense[None-shape11]" "$P::test_dense_wide_support_cell_runs_on_the_tensor_core_kernels" \
        > gpurun_out/memcheck2_$T.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|rror:" gpurun_out/memcheck2_$T.log | head -12 ;;
    *) echo "unknown stage $stage" ;;
  esac
done
ls -la gpurun_out/ | tail -12
