#!/bin/bash
# A/B of the two fused GEMM+LayerNorm kernels: isolated (cold operands) and inside the 12-layer forward
mkdir -p gpurun_out
timeout 400 python tools/lnrow_probe.py > gpurun_out/lnrow_probe.log 2>&1; echo "probe rc=$?"
sed -n '/== timing/,$p' gpurun_out/lnrow_probe.log
for cfg in "1 4" "2 4" "2 3"; do
  set -- $cfg
  MMR_GEMM_LN=$1 MMR_LN_LONGK_STAGES=$2 timeout 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_ln$1_s$2.json 2> gpurun_out/bench_ln$1_s$2.err
  echo "LN=$1 stages=$2 rc=$?"
  python - gpurun_out/bench_ln$1_s$2.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print(" value", round(d["value"]), "ms", round(d["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"], {n:(round(v["avg_launch_us"],1)) for n,v in k.items()})
PY
done
