"""Pipelined tcgen05 attention: what paces it?  Same shape with / without a key mask, cold (rotating buffers) / hot."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
B, Sq, Sk = 256, 68, 68
torch.manual_seed(5)
bufs = []
for i in range(3):
    qkv = torch.randn(B * Sq, 3 * 768, device="cuda").half()
    lens = torch.randint(1, Sk + 1, (B,), device="cuda")
    mask = (torch.arange(Sk, device="cuda")[None, :] < lens[:, None]).int().contiguous()
    bufs.append((qkv, mask))
for tc in (2, 0):
    _lib.check(lib.mmr_set_tuning(_lib.TUNE_ATTN_TC, tc))
    for masked in (True, False):
        for hot in (False, True):
            def run(i):
                a, m = bufs[0 if hot else i % 3]
                ops.attention(a[:, :768], a[:, 768:1536], a[:, 1536:], m if masked else None, B, Sq, Sk, 12)
            for i in range(6):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 60
            e0.record()
            for i in range(n):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            print(f"tc={tc} masked={masked} hot={hot}: {e0.elapsed_time(e1) * 1e3 / n:.1f} us/launch", flush=True)
_lib.check(lib.mmr_set_tuning(_lib.TUNE_ATTN_TC, 0))
