#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for pdl in 1 0; do
MMR_PDL=$pdl timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.log 2>&1; echo "bench pdl=$pdl rc=$?"
done
python - <<PY
import json
for f in ("bench_pdl1","bench_pdl0"):
    try:
        l=open(f"gpurun_out/{f}.log").read().strip().split("\n")[-1]
        d=json.loads(l); print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["share_of_step"], d["roofline"]["whole_step"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
