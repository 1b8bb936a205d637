"""Phase timeline of the fused GEMM+LayerNorm kernel (debug stamps, %globaltimer): where an epilogue warp's time goes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
lib.mmr_debug_set_ln_trace.argtypes = [C.c_void_p]
M = 17408
for K in (768, 3072):
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(768, K, device="cuda") * 0.03).half()
    b = torch.randn(768, device="cuda") * 0.1
    x = torch.randn(M, 768, device="cuda")
    g = torch.rand(768, device="cuda") + 0.5
    be = torch.randn(768, device="cuda") * 0.1
    for _ in range(3):
        ops.gemm_layernorm(a, w, b, x, g, be)
    tr = torch.zeros((144, 8, 4, 8), dtype=torch.int64, device="cuda")
    lib.mmr_debug_set_ln_trace(tr.data_ptr())
    ops.gemm_layernorm(a, w, b, x, g, be)
    torch.cuda.synchronize()
    lib.mmr_debug_set_ln_trace(None)
    t = tr.cpu().numpy().astype(np.float64)
    t0 = t[t > 0].min()
    names = ["top", "tfull", "pass1", "exch", "pass2"]
    print(f"K={K}: kernel span {(t.max() - t0) / 1e3:.1f} us")
    for it in range(3):
        v = t[:, :, it, :5]
        ok = v[:, :, 0] > 0
        rel = (v[ok] - t0) / 1e3
        d = np.diff(rel, axis=1)
        print(f"  tile {it}: n={ok.sum():4d} start(avg)={rel[:,0].mean():6.1f}us | wait_tfull {d[:,0].mean():5.1f} (max {d[:,0].max():5.1f}) | "
              f"pass1 {d[:,1].mean():5.1f} (max {d[:,1].max():5.1f}) | exchange {d[:,2].mean():5.1f} (max {d[:,2].max():5.1f}) | "
              f"pass2 {d[:,3].mean():5.1f} (max {d[:,3].max():5.1f}) | end(avg)={rel[:,4].mean():6.1f} end(max)={rel[:,4].max():6.1f}")
