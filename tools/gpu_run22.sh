#!/bin/bash
mkdir -p gpurun_out
for K in 30 100; do
MMR_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $K --warmup 5 --no-e2e > gpurun_out/bench_2gpu_$K.log 2>&1; echo "K=$K rc=$?"; grep -E "DEBUG|value" gpurun_out/bench_2gpu_$K.log | cut -c1-250
done
