"""First-contact probe for the CUDA operators on a real B200: each case runs in its own subprocess with a
timeout so that a hung kernel cannot take the whole call down.  Prints max abs / rel error per case.

    python tools/gpu_op_probe.py            # all cases
    python tools/gpu_op_probe.py gemm_small # one case (used by the parent)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _err(name, got, ref):
    import torch
    got = got.float()
    ref = ref.float()
    d = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-30
    bad = torch.isnan(got).sum().item()
    print(f"{name}: max_abs={d.max().item():.3e} mean_abs={d.mean().item():.3e} rel_to_max={d.max().item()/denom:.3e} "
          f"ref_absmax={denom:.3e} nan={bad}", flush=True)
    return d.max().item() / denom


def _gemm_case(M, N, K, dtype, act=0, residual=False, bias=True, want16=True, want32=True, lda_pad=0):
    import torch
    import torch.nn.functional as F
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    torch.manual_seed(0)
    a_full = (torch.randn(M, K + lda_pad, device="cuda") * 1.0).to(dtype)
    a = a_full[:, :K]
    w = (torch.randn(N, K, device="cuda") * 0.05).to(dtype)
    b = torch.randn(N, device="cuda") * 0.1 if bias else None
    r = torch.randn(M, N, device="cuda") if residual else None
    o16, o32 = ops.gemm(a, w, b, r, act=act, want16=want16, want32=want32)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b
    if act == 1:
        ref = F.relu(ref)
    elif act == 2:
        ref = F.gelu(ref, approximate="tanh")
    elif act == 3:
        ref = F.gelu(ref)
    elif act == 4:
        ref = torch.tanh(ref)
    if residual:
        ref = ref + r
    tag = f"gemm M={M} N={N} K={K} {str(dtype)[6:]} act={act} res={int(residual)}"
    worst = 0.0
    if o32 is not None:
        worst = max(worst, _err(tag + " out32", o32, ref))
    if o16 is not None:
        worst = max(worst, _err(tag + " out16", o16, ref))
    return worst


@case
def gemm_small():
    import torch
    return _gemm_case(128, 256, 64, torch.bfloat16, bias=False)


@case
def gemm_k768():
    import torch
    return _gemm_case(256, 768, 768, torch.bfloat16)


@case
def gemm_tail():
    import torch
    w = _gemm_case(300, 272, 128, torch.bfloat16, lda_pad=8)
    return max(w, _gemm_case(100, 16, 64, torch.float16))


@case
def gemm_acts():
    import torch
    w = 0.0
    for act in (1, 2, 3, 4):
        w = max(w, _gemm_case(384, 512, 256, torch.bfloat16, act=act))
    return max(w, _gemm_case(384, 768, 3072, torch.float16, residual=True, want16=False))


@case
def gemm_full():
    import torch
    w = _gemm_case(17408, 2304, 768, torch.bfloat16, want32=False)
    w = max(w, _gemm_case(17408, 3072, 768, torch.bfloat16, act=2, want32=False))
    return max(w, _gemm_case(17408, 768, 3072, torch.bfloat16, residual=True, want16=False))


@case
def gemm_time():
    import torch
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    for (M, N, K, act, res) in [(17408, 2304, 768, 0, False), (17408, 768, 768, 0, True), (17408, 3072, 768, 2, False),
                                (17408, 768, 3072, 0, True), (8192, 8192, 8192, 0, False)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda") if res else None
        for _ in range(3):
            ops.gemm(a, w, b, r, act=act, want16=not res, want32=res)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            ops.gemm(a, w, b, r, act=act, want16=not res, want32=res)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"gemm_time M={M} N={N} K={K} act={act} res={int(res)}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s",
              flush=True)
        # cuBLAS reference point (library GEMM, for context only)
        for _ in range(3):
            a @ w.t()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            a @ w.t()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"   cublas plain matmul: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    return 0.0


@case
def layernorm():
    import torch
    import torch.nn.functional as F
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    torch.manual_seed(1)
    w = 0.0
    for (M, H) in [(1000, 768), (33, 1536)]:
        x = torch.randn(M, H, device="cuda") * 3 + 0.5
        g = torch.rand(H, device="cuda") + 0.5
        b = torch.randn(H, device="cuda") * 0.1
        o16, o32 = ops.layernorm(x, g, b, dtype=torch.bfloat16)
        torch.cuda.synchronize()
        ref = F.layer_norm(x, (H,), g, b, 1e-12)
        w = max(w, _err(f"layernorm {M}x{H} out32", o32, ref))
        _err(f"layernorm {M}x{H} out16", o16, ref)
    return w


@case
def attention():
    import torch
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    torch.manual_seed(2)
    worst = 0.0
    for dtype in (torch.bfloat16, torch.float16):
        for (B, Sq, Sk, masked) in [(3, 68, 68, True), (2, 32, 36, True), (2, 36, 32, False), (2, 104, 104, False),
                                    (1, 128, 128, True), (2, 10, 23, True)]:
            H = 12
            qkv_q = (torch.randn(B * Sq, 3 * 768, device="cuda")).to(dtype)
            qkv_k = qkv_q if Sq == Sk else (torch.randn(B * Sk, 3 * 768, device="cuda")).to(dtype)
            mask = None
            if masked:
                lens = torch.randint(1, Sk + 1, (B,), device="cuda")
                mask = (torch.arange(Sk, device="cuda")[None, :] < lens[:, None]).int().contiguous()
            out = ops.attention(qkv_q[:, :768], qkv_k[:, 768:1536], qkv_k[:, 1536:], mask, B, Sq, Sk, H)
            torch.cuda.synchronize()
            q = qkv_q[:, :768].float().view(B, Sq, H, 64).transpose(1, 2)
            k = qkv_k[:, 768:1536].float().view(B, Sk, H, 64).transpose(1, 2)
            v = qkv_k[:, 1536:].float().view(B, Sk, H, 64).transpose(1, 2)
            s = q @ k.transpose(-1, -2) / 8.0
            if mask is not None:
                s = s + (1.0 - mask.float())[:, None, None, :] * -10000.0
            p = torch.softmax(s, -1)
            ref = (p @ v).transpose(1, 2).reshape(B * Sq, 768)
            worst = max(worst, _err(f"attention {str(dtype)[6:]} B={B} Sq={Sq} Sk={Sk} mask={int(masked)}", out, ref))
    return worst


@case
def cast():
    import torch
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    x = torch.randn(36 * 7, 2048, device="cuda")
    o = ops.cast16(x, torch.bfloat16)
    torch.cuda.synchronize()
    return _err("cast16", o, x.bfloat16())


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        worst = CASES[name]()
        print(f"CASE {name} worst_rel={worst:.3e}", flush=True)
        return
    summary = []
    for name in CASES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=240)
            out = r.stdout + ("\n[stderr]\n" + r.stderr[-3000:] if r.returncode != 0 else "")
            status = f"rc={r.returncode}"
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            status = "TIMEOUT"
        print(f"===== {name}: {status} ({time.time()-t0:.1f}s)\n{out}", flush=True)
        summary.append((name, status))
    print("SUMMARY", summary)


if __name__ == "__main__":
    main()
