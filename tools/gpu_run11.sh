#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_ops.log 2>&1
echo "pytest ops rc=$?"; tail -3 gpurun_out/pytest_ops.log
for c in 2 1; do
  echo "== MMR_GEMM_CLUSTER=$c"
  MMR_GEMM_CLUSTER=$c timeout 300 python tools/gpu_gemm_probe.py 1 2>&1 | grep -E "time M=(17408|26624|8192|9216) N=(2304|3072|8192|768) K=(768|8192|2048) act=(0|2|3|1) res=0"
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
for f in ("bench",):
    try:
        l=open(f"gpurun_out/{f}.log").read().strip().split("\n")[-1]
        d=json.loads(l); print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["share_of_step"], d["roofline"]["whole_step"])
    except Exception as e: print(f, "ERR", e)
PY
