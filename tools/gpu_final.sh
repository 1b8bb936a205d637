#!/bin/bash
# round-end check on one B200: GPU test-suite, smoke, bench lines of the three scorers (+ the CPU reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
for m in imagebert_zk imagebert_lds lxmert; do
  timeout 600 python bench.py --model $m > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "$m rc=$?"
  python - gpurun_out/bench_$m.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]
print(" value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), "clk", d["clocks"], "cpu", round(d["cpu_baseline"]["value"],1), d["cpu_baseline"]["cores"], "roofline frac", round(r["frac"],3), "whole", r.get("whole_step"))
PY
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
