#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fused_gemm_layernorm or alternate" > gpurun_out/pytest_ln.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ln.log
timeout 200 python tools/ln_trace.py 2>&1 | tail -9
for i in 1 2; do
timeout 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_ln_sw.json 2> gpurun_out/bench_ln_sw.err; echo "rc=$?"
python - gpurun_out/bench_ln_sw.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print(" value", round(d["value"]), "ms", round(d["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"], {n:(v["launches_per_step"], round(v["avg_launch_us"],1)) for n,v in k.items()})
PY
done
