#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "fused or alternate" > gpurun_out/pytest_ops.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ops.log
timeout 200 python tools/ln_trace.py 2>&1 | tail -8
timeout 300 python tools/gpu_gemm_probe.py ln 2>&1 | grep "time"
