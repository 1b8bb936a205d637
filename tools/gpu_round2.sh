#!/bin/bash
# round-2 check on one B200: GPU test-suite, smoke, the full bench line (headline + other models + strict + cfg4 / cfg5)
# and the CPU reference arm.  Output under gpurun_out/ with the given tag (default r02).
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1300 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/${TAG}_pytest.log | tail -8
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/${TAG}_bench.err
python - gpurun_out/${TAG}_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]), d["clocks"], round(d["roofline"]["frac"], 3),
      d["roofline"]["whole_step"])
for k, v in (d["other_models"] or {}).items():
    print(k, round(v["value"]), round(v["frac_of_peak"], 3), v["peak"], v["launches_per_forward"])
print(d["strict"]); print(d["cfg4"]); print(d["cfg5"]); print(d["cpu_baseline"])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
# ncu launch list of the same bench command (every launch; the summary keeps the last two forwards = the step's kernels)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_under_ncu.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv -138 | tee gpurun_out/${TAG}_launches_step.txt
