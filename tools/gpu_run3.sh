#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/gpu_gemm_probe.py 1 > gpurun_out/gemm_probe.log 2>&1; tail -9 gpurun_out/gemm_probe.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
