#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
l=open("gpurun_out/bench.log").read().strip().split("\n")[-1]
d=json.loads(l); print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["achieved"], {k:(round(v["avg_launch_us"],1), round(v["share_of_step"],3)) for k,v in d["roofline"]["kernels"].items()}, d["roofline"]["whole_step"], d["clocks"])
PY
