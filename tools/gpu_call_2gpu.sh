#!/bin/bash
# 2-GPU session: NCCL halo test, data-parallel bench line (weak scaling), reference arm under torchrun
mkdir -p gpurun_out
T=${1:-r1q}
nvidia-smi --query-gpu=name --format=csv,noheader | tee gpurun_out/gpus_$T.txt
timeout 300 python -m pytest tests/test_halo_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_2gpu_$T.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench_2gpu_$T.err > gpurun_out/bench_2gpu_$T.json; tail -2 gpurun_out/bench_2gpu_$T.err; cut -c1-400 gpurun_out/bench_2gpu_$T.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 2> gpurun_out/ref_2gpu_$T.err > gpurun_out/ref_2gpu_$T.json; cut -c1-200 gpurun_out/ref_2gpu_$T.json
