"""Per-item timeline of the pipelined tcgen05 attention kernel (attention_tc2.cu debug stamps, %globaltimer)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
lib.mmr_debug_set_attn_trace.argtypes = [C.c_void_p]
_lib.check(lib.mmr_set_tuning(_lib.TUNE_ATTN_TC, 2))
B, Sq, Sk = 256, 68, 68
torch.manual_seed(1)
qkv = torch.randn(B * Sq, 3 * 768, device="cuda").half()
lens = torch.randint(1, Sk + 1, (B,), device="cuda")
mask = (torch.arange(Sk, device="cuda")[None, :] < lens[:, None]).int().contiguous()
for _ in range(3):
    ops.attention(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], mask, B, Sq, Sk, 12)
tr = torch.zeros((148, 24, 16), dtype=torch.int64, device="cuda")
lib.mmr_debug_set_attn_trace(tr.data_ptr())
ops.attention(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], mask, B, Sq, Sk, 12)
torch.cuda.synchronize()
lib.mmr_debug_set_attn_trace(None)
t = tr.cpu().numpy().astype(np.float64)
t0 = t[t > 0].min()
print(f"span {(t.max() - t0) / 1e3:.1f} us")
names = ["tma", "qk", "pv", "s_seen", "p_arr", "o_seen", "done"]
for cta in (0, 77):
    print(f"CTA {cta}")
    for n in range(24):
        if t[cta, n, 0] == 0:
            break
        r = (t[cta, n, :7] - t0) / 1e3
        pr = (t[cta, n, 7:12] - t0) / 1e3
        print(f"  item {n:2d}: " + " ".join(f"{names[i]} {r[i]:6.2f}" for i in range(7)) +
              f" | p_arr w1 {pr[0]:6.2f} w2 {pr[1]:6.2f} | mma: p_ready seen {pr[2]:6.2f} pv issued {pr[3]:6.2f} qk ready seen {pr[4]:6.2f}")
v = t[:, :21, :7]
ok = v[:, :, 0] > 0
r = (v - t0) / 1e3
print("mean over CTAs, per item: qk-tma, s_seen-qk, p_arr-s_seen, pv-p_arr, o_seen-pv, done-o_seen")
for n in range(21):
    m = ok[:, n]
    if m.sum() == 0:
        break
    x = r[m, n]
    print(f"  item {n:2d}: {np.mean(x[:,1]-x[:,0]):6.2f} {np.mean(x[:,3]-x[:,1]):6.2f} {np.mean(x[:,4]-x[:,3]):6.2f} "
          f"{np.mean(x[:,2]-x[:,4]):6.2f} {np.mean(x[:,5]-x[:,2]):6.2f} {np.mean(x[:,6]-x[:,5]):6.2f}   done at {np.mean(x[:,6]):6.2f}")
