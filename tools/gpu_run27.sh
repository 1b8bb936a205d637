#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_parity.log
for d in 0 1 0 1; do
  MMR_LABEL_DEDUP=$d timeout 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_dedup$d.json 2> gpurun_out/bench_dedup$d.err
  echo "DEDUP=$d rc=$?"
  python - gpurun_out/bench_dedup$d.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print(" value", round(d["value"]), "ms", round(d["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"], {n:(v["launches_per_step"], round(v["avg_launch_us"],1)) for n,v in k.items()})
PY
done
