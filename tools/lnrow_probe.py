"""Row-owner GEMM+LayerNorm kernel (gemm_lnrow_sm100.cu) against the three-pairs-per-block one (gemm_ln_sm100.cu):
correctness against the unfused fp32 reference, launch time of both (CUDA events, L2-cold rotating operands), and
the phase timeline of the new kernel's epilogue warps (%globaltimer stamps)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
lib.mmr_debug_set_lnrow_trace.argtypes = [C.c_void_p]


def rel(a, b):
    return ((a.float() - b).abs().max() / b.abs().max()).item()


def make(M, K, seed):
    torch.manual_seed(seed)
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(768, K, device="cuda") * 0.03).half()
    b = torch.randn(768, device="cuda") * 0.1
    x = torch.randn(M, 768, device="cuda") * 2 + 0.3
    g = torch.rand(768, device="cuda") + 0.5
    be = torch.randn(768, device="cuda") * 0.1
    return a, w, b, x, g, be


def set_ln(v, cfg=0):
    _lib.check(lib.mmr_set_tuning(_lib.TUNE_GEMM_LN, v))
    _lib.check(lib.mmr_set_tuning(_lib.TUNE_LN_ROW_CFG, cfg))


print("== correctness (new kernel) ==")
CFGS = (421, 331, 511, 412, 322)
for stages in CFGS:
    set_ln(2, stages)
    for (M, K) in [(17408, 768), (17408, 3072), (300, 768), (1000, 3072), (129, 768), (38000, 768)]:
        a, w, b, x, g, be = make(M, K, M + K)
        ref = F.layer_norm(a.float() @ w.float().t() + b + x, (768,), g, be, 1e-12)
        x16, x32 = ops.gemm_layernorm(a, w, b, x, g, be)
        torch.cuda.synchronize()
        print(f"stages={stages} M={M} K={K}: rel32 {rel(x32, ref):.2e} rel16 {rel(x16, ref):.2e}", flush=True)

print("== timing ==")
for (M, K) in [(17408, 768), (17408, 3072), (26624, 768)]:
    sets = [make(M, K, 7 + i) for i in range(4)]   # > L2 in total for the big shapes
    for label, v, stages in [("old", 1, 0)] + [(f"row/{c}", 2, c) for c in CFGS]:
        set_ln(v, stages)
        for s in sets:
            ops.gemm_layernorm(*s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 40
        e0.record()
        for i in range(n):
            ops.gemm_layernorm(*sets[i % 4])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        print(f"M={M} K={K} {label}: {us:.1f} us/launch ({2.0 * M * 768 * K / us / 1e6:.0f} TFLOP/s)", flush=True)

print("== timeline (new kernel) ==")
names = ["start", "tfull0", "p1end0", "tfull1", "p1end1", "tfull2", "p1end2", "stats", "p2a", "p2b"]
idx = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
for (M, K, stages) in [(17408, 768, 421), (17408, 768, 511), (17408, 768, 322), (17408, 3072, 511)]:
    set_ln(2, stages)
    a, w, b, x, g, be = make(M, K, 3)
    for _ in range(3):
        ops.gemm_layernorm(a, w, b, x, g, be)
    tr = torch.zeros((148, 8, 2, 16), dtype=torch.int64, device="cuda")
    lib.mmr_debug_set_lnrow_trace(tr.data_ptr())
    ops.gemm_layernorm(a, w, b, x, g, be)
    torch.cuda.synchronize()
    lib.mmr_debug_set_lnrow_trace(None)
    t = tr.cpu().numpy().astype(np.float64)
    t0 = t[t > 0].min()
    print(f"M={M} K={K} stages={stages}: span {(t.max() - t0) / 1e3:.1f} us")
    for blk in range(2):
        v = t[:, :, blk, :10]
        ok = v[:, :, 0] > 0
        if ok.sum() == 0:
            continue
        r = (v[ok] - t0) / 1e3
        print(f"  block {blk} (n={ok.sum()}): " + " | ".join(f"{names[i]} {r[:, i].mean():5.1f}/{r[:, i].max():5.1f}" for i in idx))
set_ln(1)
