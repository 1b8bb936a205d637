#!/bin/bash
# round-end check on one B200: GPU test-suite, smoke, bench lines of the three scorers (+ the CPU reference arm),
# ncu launch list of the bench steps
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
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 138 -c 140 --csv --log-file gpurun_out/launches_bench_zk_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_bench_zk_cfg2.csv > gpurun_out/launches_bench_zk_cfg2.txt; cat gpurun_out/launches_bench_zk_cfg2.txt
