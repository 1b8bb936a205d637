#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/attn_trace.py > gpurun_out/attn_trace3.log 2>&1; echo "trace rc=$?"; sed -n 1,2p gpurun_out/attn_trace3.log; tail -22 gpurun_out/attn_trace3.log
timeout 300 python tools/attn_probe.py tcgen05_pipelined mma_sync 2>&1 | grep -E "us/launch|worst|FAIL"
for tc in 0 2; do
  MMR_ATTN_TC=$tc timeout 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_tc$tc.json 2> gpurun_out/bench_tc$tc.err
  echo "ATTN_TC=$tc rc=$?"
  python - gpurun_out/bench_tc$tc.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print(" value", round(d["value"]), "ms", round(d["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"], {n:(round(v["avg_launch_us"],1)) for n,v in k.items()})
PY
done
