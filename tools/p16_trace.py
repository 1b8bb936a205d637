"""Timeline of the 16-bit-output GEMM kernel (debug stamps): prologue, per-tile MMA issue window, per-tile epilogue,
teardown -- for one launch of the QKV and FFN-in shapes, with the previous launch still draining (back-to-back)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
lib.mmr_debug_set_p16_trace.argtypes = [C.c_void_p]
M = 17408
for (N, K, act) in [(2304, 768, 0), (3072, 768, 2)]:
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") * 0.05).half()
    b = torch.randn(N, device="cuda")
    for _ in range(3):
        ops.gemm(a, w, b, None, act=act)
    tr = torch.zeros((74, 128), dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.gemm(a, w, b, None, act=act)
    lib.mmr_debug_set_p16_trace(tr.data_ptr())
    e0.record()
    ops.gemm(a, w, b, None, act=act)
    e1.record()
    lib.mmr_debug_set_p16_trace(None)
    ops.gemm(a, w, b, None, act=act)
    torch.cuda.synchronize()
    t = tr.cpu().numpy().astype(np.float64)
    t0 = t[:, 0].min()
    u = lambda x: (x - t0) / 1e3
    print(f"N={N} K={K}: event time {e0.elapsed_time(e1) * 1e3:.1f} us; first entry -> last exit {u(t[:, 2].max()):.1f} us")
    print(f"  entry spread {u(t[:, 0]).max():.1f} us | prologue (entry -> roles) avg {np.mean(t[:, 1] - t[:, 0]) / 1e3:.2f} us | "
          f"exit: first {u(t[:, 2].min()):.1f} last {u(t[:, 2].max()):.1f}")
    tiles = t[:, 4:].reshape(74, 31, 4)
    for it in range(10):
        v = tiles[:, it, :]
        ok = v[:, 0] > 0
        if not ok.any():
            break
        v = v[ok]
        print(f"  tile {it}: n={ok.sum():2d} mma_begin {u(v[:, 0]).mean():6.1f} | issue window {np.mean(v[:, 1] - v[:, 0]) / 1e3:5.2f} | "
              f"acc ready (epi start) {u(v[:, 2]).mean():6.1f} | epilogue {np.mean(v[:, 3] - v[:, 2]) / 1e3:5.2f} (max {np.max(v[:, 3] - v[:, 2]) / 1e3:5.2f}) | "
              f"tile period {'' if it == 0 else format(np.mean(v[:, 2]) / 1e3 - prev, '5.2f')}")
        prev = np.mean(v[:, 2]) / 1e3
