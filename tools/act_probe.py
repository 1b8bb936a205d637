import torch, sys
sys.path.insert(0, ".")
from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
M, N, K = 17408, 3072, 768
a = torch.randn(M, K, device="cuda").half()
w = (torch.randn(N, K, device="cuda") * 0.05).half()
b = torch.randn(N, device="cuda")
for act in (0, 1, 2, 3, 0, 2):
    for _ in range(3): ops.gemm(a, w, b, None, act=act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): ops.gemm(a, w, b, None, act=act)
    e1.record(); torch.cuda.synchronize()
    print(f"act={act}: {e0.elapsed_time(e1)/30*1e3:.1f} us")
