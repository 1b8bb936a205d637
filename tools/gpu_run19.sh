#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "attention" > gpurun_out/pytest_ops.log 2>&1
echo "pytest attention rc=$?"; tail -25 gpurun_out/pytest_ops.log | cut -c1-200
timeout 120 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops, _lib
lib = _lib.load()
B, S = 256, 68
qkv = torch.randn(B * S, 2304, device="cuda").half()
mask = torch.ones(B, S, dtype=torch.int32, device="cuda")
for tc in (1, 0):
    lib.mmr_set_tuning(_lib.TUNE_ATTN_TC, tc)
    for _ in range(3):
        ops.attention(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], mask, B, S, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attention(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], mask, B, S, S)
    e1.record(); torch.cuda.synchronize()
    print(f"attention B=256 S=68 tcgen05={tc}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
PY
