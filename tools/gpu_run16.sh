#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k attention > gpurun_out/pytest_ops.log 2>&1
echo "pytest attention rc=$?"; tail -5 gpurun_out/pytest_ops.log
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops, _lib
lib = _lib.load()
B, S = 256, 68
qkv = torch.randn(B * S, 2304, device="cuda").half()
mask = torch.ones(B, S, dtype=torch.int32, device="cuda")
for tma in (1, 0):
    lib.mmr_set_tuning(_lib.TUNE_ATTN_TMA, tma)
    for _ in range(3):
        ops.attention(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], mask, B, S, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attention(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], mask, B, S, S)
    e1.record(); torch.cuda.synchronize()
    print(f"attention B=256 S=68 tma={tma}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
PY
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
l=open("gpurun_out/bench.log").read().strip().split("\n")[-1]
d=json.loads(l); print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["achieved"], {k:(round(v["avg_launch_us"],1), round(v["share_of_step"],3)) for k,v in d["roofline"]["kernels"].items()}, d["roofline"]["whole_step"])
PY
