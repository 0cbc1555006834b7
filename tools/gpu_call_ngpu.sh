#!/bin/bash
# N-GPU data-parallel bench line (weak scaling): bash tools/gpu_call_ngpu.sh TAG N
mkdir -p gpurun_out
T=${1:-r1q}; N=${2:-4}
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_${N}gpu_$T.err > gpurun_out/bench_${N}gpu_$T.json; tail -2 gpurun_out/bench_${N}gpu_$T.err; cut -c1-260 gpurun_out/bench_${N}gpu_$T.json
