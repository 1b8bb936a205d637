#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1; echo "bench 2gpu rc=$?"; tail -2 gpurun_out/bench_2gpu.log | cut -c1-600
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_1gpu.log 2>&1; echo "bench 1gpu rc=$?"; tail -1 gpurun_out/bench_1gpu.log | cut -c1-3000
