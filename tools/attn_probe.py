"""Attention kernels side by side: correctness of the selected variant on the test shapes, launch time of every
variant at the bench shapes (CUDA events, rotating qkv buffers larger than L2 together).  One Python call per launch
costs ~20 us of host time (tensor maps, ctypes): below that the numbers are host-bound — tools/attn_ablate.py replays
the launches from a CUDA graph for GPU-bound times."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
VARIANTS = {"mma_sync": (0, 0), "mma_sync_tma": (0, 1), "tcgen05": (1, 0), "tcgen05_pipelined": (2, 0)}


def select(name):
    tc, tma = VARIANTS[name]
    _lib.check(lib.mmr_set_tuning(_lib.TUNE_ATTN_TC, tc))
    _lib.check(lib.mmr_set_tuning(_lib.TUNE_ATTN_TMA, tma))


def reference(qkv_q, qkv_k, mask, B, Sq, Sk, H=12):
    q = qkv_q[:, :768].float().view(B, Sq, H, 64).transpose(1, 2)
    k = qkv_k[:, 768:1536].float().view(B, Sk, H, 64).transpose(1, 2)
    v = qkv_k[:, 1536:].float().view(B, Sk, H, 64).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / 8.0
    if mask is not None:
        s = s + (1.0 - mask.float())[:, None, None, :] * -10000.0
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * Sq, 768)


only = sys.argv[1:] or list(VARIANTS)
print("== correctness ==")
for name in only:
    select(name)
    torch.manual_seed(2)
    worst = 0.0
    for dtype in (torch.float16, torch.bfloat16):
        for (B, Sq, Sk, masked) in [(3, 68, 68, True), (2, 32, 36, True), (2, 36, 32, False), (2, 104, 104, False),
                                    (1, 128, 128, True), (2, 10, 23, True), (2, 1, 1, False), (2, 28, 28, True),
                                    (256, 68, 68, True), (300, 36, 32, True)]:
            qkv_q = torch.randn(B * Sq, 3 * 768, device="cuda").to(dtype)
            qkv_k = qkv_q if Sq == Sk else torch.randn(B * Sk, 3 * 768, device="cuda").to(dtype)
            mask = None
            if masked:
                lens = torch.randint(1, Sk + 1, (B,), device="cuda")
                mask = (torch.arange(Sk, device="cuda")[None, :] < lens[:, None]).int().contiguous()
            out = ops.attention(qkv_q[:, :768], qkv_k[:, 768:1536], qkv_k[:, 1536:], mask, B, Sq, Sk, 12)
            torch.cuda.synchronize()
            ref = reference(qkv_q, qkv_k, mask, B, Sq, Sk)
            err = ((out.float() - ref).abs().max() / ref.abs().max()).item()
            tol = 5e-3 if dtype == torch.bfloat16 else 1.5e-3
            flag = "" if err < tol else "  <-- FAIL"
            if flag or B >= 256:
                print(f"{name} {str(dtype)[6:]} B={B} Sq={Sq} Sk={Sk} masked={masked}: rel {err:.2e}{flag}", flush=True)
            worst = max(worst, err / tol)
    print(f"{name}: worst err / tol = {worst:.2f}", flush=True)

print("== timing ==")
for (B, Sq, Sk) in [(256, 68, 68), (256, 104, 104), (256, 32, 36), (256, 36, 32), (256, 32, 32)]:
    torch.manual_seed(5)
    bufs = []
    for i in range(3):
        qkv_q = torch.randn(B * Sq, 3 * 768, device="cuda").half()
        qkv_k = qkv_q if Sq == Sk else torch.randn(B * Sk, 3 * 768, device="cuda").half()
        lens = torch.randint(1, Sk + 1, (B,), device="cuda")
        mask = (torch.arange(Sk, device="cuda")[None, :] < lens[:, None]).int().contiguous()
        bufs.append((qkv_q, qkv_k, mask))
    for name in only:
        select(name)
        for (a, b, m) in bufs:
            ops.attention(a[:, :768], b[:, 768:1536], b[:, 1536:], m, B, Sq, Sk, 12)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 60
        e0.record()
        for i in range(n):
            a, b, m = bufs[i % 3]
            ops.attention(a[:, :768], b[:, 768:1536], b[:, 1536:], m, B, Sq, Sk, 12)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        mb = (B * Sq * 768 * 2 * 2 + 2 * B * Sk * 768 * 2) / 1e6
        print(f"B={B} Sq={Sq} Sk={Sk} {name}: {us:.1f} us/launch ({mb / us * 1e-3 * 1e3:.0f} GB/s algorithmic)", flush=True)
select("mma_sync")
