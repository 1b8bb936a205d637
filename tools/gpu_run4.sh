#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "dscore|passed|failed" gpurun_out/pytest_gpu.log | tail -30
for m in imagebert_lds lxmert; do
timeout 600 python bench.py --steps 30 --warmup 5 --model $m --no-cpu-baseline > gpurun_out/bench_$m.log 2>&1; echo "bench $m rc=$?"; tail -1 gpurun_out/bench_$m.log | cut -c1-400
python - <<PY
import json
l=open("gpurun_out/bench_$m.log").read().strip().split("\n")[-1]
d=json.loads(l); print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["achieved"], d["roofline"]["share_of_step"], d["roofline"]["whole_step"])
PY
done
