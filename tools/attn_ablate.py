"""Timing-only ablations of the pipelined tcgen05 attention kernel (results are WRONG when a bit is set): which
per-item cost is the period?  bit 1: no output stores, 2: no key_mask fetch, 4: no softmax math, 8: one TMA box per
item instead of three, 16: no O read-out, 32: one MMA per batch instead of 4 / SkP/16."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
lib.mmr_debug_set_attn_ablate.argtypes = [C.c_int]
for (B, Sq, Sk) in [(256, 68, 68), (256, 32, 32)]:
    torch.manual_seed(5)
    bufs = []
    for i in range(3):
        qkv = torch.randn(B * Sq, 3 * 768, device="cuda").half()
        lens = torch.randint(1, Sk + 1, (B,), device="cuda")
        mask = (torch.arange(Sk, device="cuda")[None, :] < lens[:, None]).int().contiguous()
        bufs.append((qkv, mask))
    for bits in (0, 1, 2, 4, 8, 16, 32, 1 | 16, 4 | 16, 31, 63):
        lib.mmr_debug_set_attn_ablate(bits)
        outs = [torch.empty((B * Sq, 768), dtype=torch.float16, device="cuda") for _ in range(3)]

        def run(i):
            a, m = bufs[i % 3]
            _lib.check(lib.mmr_attention(a[:, :768].data_ptr(), 2304, a[:, 768:1536].data_ptr(), 2304,
                                         a[:, 1536:].data_ptr(), 2304, m.data_ptr(), outs[i % 3].data_ptr(), 768, B, Sq,
                                         Sk, 12, _lib.DT_FP16, torch.cuda.current_stream().cuda_stream))
        s_ = torch.cuda.Stream()
        with torch.cuda.stream(s_):
            for i in range(6):
                run(i)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s_):      # 30 launches back to back: GPU-bound, no per-call host time
                for i in range(30):
                    run(i)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
        print(f"S={Sq}x{Sk} ablate={bits:2d}: {e0.elapsed_time(e1) * 1e3 / 150:.1f} us/launch (CUDA graph of 30)", flush=True)
lib.mmr_debug_set_attn_ablate(0)
