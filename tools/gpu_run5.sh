#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/gpu_gemm_probe.py ln > gpurun_out/gemm_ln_probe.log 2>&1; tail -14 gpurun_out/gemm_ln_probe.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
python - <<PY
import json
l=open("gpurun_out/bench.log").read().strip().split("\n")[-1]
d=json.loads(l); print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["share_of_step"], d["roofline"]["whole_step"])
PY
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
