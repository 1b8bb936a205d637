#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/lnrow_probe.py > gpurun_out/lnrow_probe2.log 2>&1; echo "probe rc=$?"
grep -c "rel32" gpurun_out/lnrow_probe2.log; grep rel32 gpurun_out/lnrow_probe2.log | awk '{print $4, $6}' | sort -g | tail -2
sed -n '/== timing/,$p' gpurun_out/lnrow_probe2.log
for cfg in "1 0" "3 421" "3 511" "3 322" "2 511"; do
  set -- $cfg
  MMR_GEMM_LN=$1 MMR_LN_ROW_CFG=$2 timeout 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_ln$1_c$2.json 2> gpurun_out/bench_ln$1_c$2.err
  echo "LN=$1 cfg=$2 rc=$?"
  python - gpurun_out/bench_ln$1_c$2.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print(" value", round(d["value"]), "ms", round(d["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"], {n:(round(v["avg_launch_us"],1)) for n,v in k.items()})
PY
done
