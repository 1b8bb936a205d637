#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_ops.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/pytest_ops.log
timeout 400 python tools/gpu_gemm_probe.py ln > gpurun_out/gemm_ln_probe.log 2>&1; grep time gpurun_out/gemm_ln_probe.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
for f in ("bench",):
    try:
        l=open(f"gpurun_out/{f}.log").read().strip().split("\n")[-1]
        d=json.loads(l); print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["share_of_step"], d["roofline"]["whole_step"])
    except Exception as e: print(f, "ERR", e)
PY
