"""GPU probe of the tcgen05 GEMM kernels: correctness on the encoder shapes and device time per shape, for the
CTA-pair kernel and (MMR_GEMM_PAIR=0) the single-CTA kernel; cuBLAS time beside it for context only.
Each mode runs in a subprocess with a timeout so a hung pipeline cannot take the whole gpurun call down."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [(17408, 2304, 768, 0, False), (17408, 768, 768, 0, True), (17408, 3072, 768, 2, False),
          (17408, 768, 3072, 0, True), (26624, 3072, 768, 2, False), (8192, 3072, 768, 3, False),
          (9216, 768, 2048, 1, False), (8192, 8192, 8192, 0, False)]


def child():
    import torch
    import torch.nn.functional as F
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    torch.manual_seed(0)
    acts = {0: lambda x: x, 1: F.relu, 2: lambda x: F.gelu(x, approximate="tanh"), 3: F.gelu, 4: torch.tanh}
    # correctness (incl. M tails and odd tile counts)
    for (M, N, K, act, res) in [(300, 256, 128, 0, False), (1000, 768, 768, 2, True), (17408, 768, 768, 0, True),
                                (4500, 2304, 768, 0, False), (17408, 3072, 768, 2, False), (17408, 768, 3072, 0, True)]:
        a = torch.randn(M, K, device="cuda").half()
        w = (torch.randn(N, K, device="cuda") * 0.05).half()
        b = torch.randn(N, device="cuda") * 0.1
        r = torch.randn(M, N, device="cuda") if res else None
        o16, o32 = ops.gemm(a, w, b, r, act=act, want16=True, want32=True)
        torch.cuda.synchronize()
        ref = acts[act](a.float() @ w.float().t() + b)
        if res:
            ref = ref + r
        e32 = ((o32 - ref).abs().max() / ref.abs().max()).item()
        e16 = ((o16.float() - ref).abs().max() / ref.abs().max()).item()
        print(f"check M={M} N={N} K={K} act={act} res={int(res)}: rel32={e32:.2e} rel16={e16:.2e}", flush=True)
        assert e32 < 2e-5 and e16 < 1e-3
    for (M, N, K, act, res) in SHAPES:
        a = torch.randn(M, K, device="cuda").half()
        w = (torch.randn(N, K, device="cuda") * 0.05).half()
        b = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda") if res else None
        o32 = r.clone() if res else None
        for _ in range(3):
            ops.gemm(a, w, b, r, act=act, want16=not res, want32=res)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            ops.gemm(a, w, b, r, act=act, want16=not res, want32=res)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        for _ in range(3):
            a @ w.t()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            a @ w.t()
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / n
        print(f"time M={M} N={N} K={K} act={act} res={int(res)}: {ms*1e3:7.1f} us {2*M*N*K/ms/1e9:7.1f} TFLOP/s   "
              f"(cuBLAS plain fp16 matmul {ms2*1e3:7.1f} us {2*M*N*K/ms2/1e9:7.1f} TFLOP/s)", flush=True)


def child_ln():
    import torch
    import torch.nn.functional as F
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    torch.manual_seed(0)
    for (M, K) in [(300, 768), (17408, 768), (17408, 3072), (26624, 768), (8192, 3072)]:
        a = torch.randn(M, K, device="cuda").half()
        w = (torch.randn(768, K, device="cuda") * 0.03).half()
        b = torch.randn(768, device="cuda") * 0.1
        x = torch.randn(M, 768, device="cuda")
        g = torch.rand(768, device="cuda") + 0.5
        be = torch.randn(768, device="cuda") * 0.1
        ref = F.layer_norm(a.float() @ w.float().t() + b + x, (768,), g, be, 1e-12)
        x16, x32 = ops.gemm_layernorm(a, w, b, x.clone(), g, be)
        torch.cuda.synchronize()
        e32 = ((x32 - ref).abs().max() / ref.abs().max()).item()
        print(f"fused-LN check M={M} K={K}: rel32={e32:.2e}", flush=True)
        assert e32 < 1e-5
        xs = x.clone()
        for _ in range(3):
            ops.gemm_layernorm(a, w, b, xs, g, be)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            ops.gemm_layernorm(a, w, b, xs, g, be)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        # unfused: GEMM(+residual, fp32 out) then LayerNorm
        for _ in range(3):
            _, y = ops.gemm(a, w, b, xs, want16=False, want32=True)
            ops.layernorm(y, g, be, dtype=torch.float16)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            _, y = ops.gemm(a, w, b, xs, want16=False, want32=True)
            ops.layernorm(y, g, be, dtype=torch.float16)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / n
        print(f"fused-LN time M={M} K={K}: {ms*1e3:7.1f} us {2*M*768*K/ms/1e9:7.1f} TFLOP/s   (GEMM+residual then LN: "
              f"{ms2*1e3:7.1f} us)", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child_ln":
        child_ln()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ln":
        try:
            r = subprocess.run([sys.executable, __file__, "child_ln"], capture_output=True, text=True, timeout=300)
            print(r.stdout + (("\n[stderr]\n" + r.stderr[-3000:]) if r.returncode else ""), f"rc={r.returncode}", flush=True)
        except subprocess.TimeoutExpired as e:
            print("TIMEOUT", (e.stdout or b"")[-3000:], flush=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    modes = [a for a in sys.argv[1:] if a in ("0", "1")] or ["1", "0"]
    for pair in modes:
        env = dict(os.environ, MMR_GEMM_PAIR=pair)
        print(f"===== MMR_GEMM_PAIR={pair}", flush=True)
        try:
            r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=300)
            print(r.stdout + (("\n[stderr]\n" + r.stderr[-3000:]) if r.returncode else ""), f"rc={r.returncode}", flush=True)
        except subprocess.TimeoutExpired as e:
            print("TIMEOUT", (e.stdout or b"")[-3000:], flush=True)
