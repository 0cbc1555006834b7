#!/bin/bash
# 2-GPU session: NCCL halo test (forward + backward), config 4 partitioned cell forward / training step at 2 GPUs
mkdir -p gpurun_out
T=${1:-r2d}
timeout 300 python -m pytest tests/test_halo_gpu.py -m gpu -x -q -k two_ranks > gpurun_out/pytest_2gpu_$T.log 2>&1; tail -1 gpurun_out/pytest_2gpu_$T.log; grep -E "^E  |FAILED|Error" gpurun_out/pytest_2gpu_$T.log | head -12
for mode in "" "--train"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bench_halo.py 2 $mode 2>> gpurun_out/halo_2gpu_$T.err | tee -a gpurun_out/halo_2gpu_$T.jsonl
done
tail -3 gpurun_out/halo_2gpu_$T.err
