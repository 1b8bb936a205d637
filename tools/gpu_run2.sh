#!/bin/bash
# GEMM probe (pair kernel), parity tests, short bench.
mkdir -p gpurun_out
timeout 400 python tools/gpu_gemm_probe.py 1 > gpurun_out/gemm_probe.log 2>&1; tail -16 gpurun_out/gemm_probe.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
