#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "16bit or fused" > gpurun_out/pytest_ops.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/pytest_ops.log
for s in 3 4 5 6; do
  echo "== stages $s"
  MMR_P16_STAGES=$s timeout 300 python tools/gpu_gemm_probe.py 1 2>&1 | grep -E "time M=(17408|26624|8192) N=(2304|3072|8192)"
done
