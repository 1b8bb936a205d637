#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_records.py tests/test_gpu_reference_api.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_records.log 2>&1
echo "pytest records rc=$?"; tail -12 gpurun_out/pytest_records.log | cut -c1-220
python tools/decode_bench.py 4096 > gpurun_out/decode_bench.log 2>&1; cat gpurun_out/decode_bench.log
