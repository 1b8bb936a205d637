"""Launches each encoder GEMM shape (and the fused GEMM+LayerNorm) a few times: the command ncu wraps for the
per-kernel `--set full` captures under profiles/.  Not a benchmark (numbers under a profiler are never bench values)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops  # noqa: E402

M = 17408
torch.manual_seed(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2] if len(sys.argv) > 2 else "all"
for (N, K, act, res) in [(2304, 768, 0, False), (768, 768, 0, True), (3072, 768, 2, False), (768, 3072, 0, True)]:
    if which not in ("all", "gemm"):
        break
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") * 0.05).half()
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if res else None
    for _ in range(reps):
        ops.gemm(a, w, b, r, act=act, want16=not res, want32=res)
    torch.cuda.synchronize()
if which in ("all", "ln"):
    for K in (768, 3072):
        a = torch.randn(M, K, device="cuda").half()
        w = (torch.randn(768, K, device="cuda") * 0.03).half()
        b = torch.randn(768, device="cuda") * 0.1
        x = torch.randn(M, 768, device="cuda")
        g = torch.rand(768, device="cuda") + 0.5
        be = torch.randn(768, device="cuda") * 0.1
        for _ in range(reps):
            ops.gemm_layernorm(a, w, b, x, g, be)
        torch.cuda.synchronize()
print("done")
