#!/bin/bash
# full GPU test-suite with the pipelined tcgen05 attention as default + bench lines of the three scorers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for m in imagebert_zk imagebert_lds lxmert; do
  timeout 400 python bench.py --model $m --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "$m rc=$?"
  python - gpurun_out/bench_$m.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print(" value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"], {n:(round(v["avg_launch_us"],1)) for n,v in k.items()})
PY
done
