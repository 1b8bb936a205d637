"""Round-2 GPU parity tests (run with -m gpu on a B200), all through the C ABI:

  * the AM-softmax margin branch (model_triple.py:56-86) at operator level and inside the full zk model, with
    straddlers of the 0.35 threshold reported explicitly;
  * the headline configuration itself (B = 256, full depth) and a cfg4-shaped candidate set (50 queries x 30
    candidates through the chunked host path) against the fp32 oracle, with a tie-aware top-5 check;
  * strict precision (two-term split operands): the "trained-like" weight sets inside the stated 1e-3;
  * the pruned last block ([CLS] rows only) against the full one;
  * several handles on one device (exchange tables are per handle; forwards on different streams are serialised).
"""
import ast
import os

import numpy as np
import pytest
import torch

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import synth
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LDS, LXMERT, ZK, ModelConfig

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3            # the stated tolerance (BASELINE.json north_star)
MARGIN, SCALE = 0.35, 30.0


def _scorer(cfg, w, max_batch, **kw):
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import MatchScorer
    return MatchScorer(cfg, w, device=0, max_batch=max_batch, **kw)


def _oracle(cfg, w, inp):
    from oracle import imagebert, lxmert
    wt, it = imagebert.to_torch(w), imagebert.to_torch(inp)
    if cfg.kind == ZK:
        return imagebert.zk_forward(wt, it, cfg.n_layers)
    if cfg.kind == LDS:
        return imagebert.lds_forward(wt, it, cfg.n_layers)
    return lxmert.forward(wt, it, cfg.n_layers, cfg.n_r_layers, cfg.n_x_layers)


def _gpu(sc, inp):
    feeds = {k: v.cuda() for k, v in sc.to_feeds(inp).items()}
    B = feeds["query_ids"].shape[0]
    pooled = torch.empty((B, sc.cfg.hidden), dtype=torch.float32, device="cuda")
    probs = sc.forward_device(feeds, pooled_out=pooled)
    torch.cuda.synchronize()
    return probs.cpu(), pooled.cpu()


def _full_cfg(kind, **kw):
    if kind == LXMERT:
        return ModelConfig(LXMERT, n_layers=9, n_r_layers=5, n_x_layers=5, lq=32, nbox=36, **kw)
    return ModelConfig(kind, n_layers=12, lq=32, nbox=36, **kw)


def _cosines(pooled, am_kernel):
    x = pooled.double()
    x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-6)
    k = torch.from_numpy(np.asarray(am_kernel)).double()
    k = k / k.norm(dim=0, keepdim=True).clamp_min(1e-5)
    return (x @ k).clamp(-1, 1)


def _aligned_am_kernel(pooled_ref: torch.Tensor) -> np.ndarray:
    """An am_kernel [768, 2] whose cosines with THESE pooled rows straddle the 0.35 margin threshold: column 1 points
    0.35 of the way along the mean pooled direction and otherwise along the direction in which the rows differ most,
    column 0 is its mirror image; the reference initialiser (Xavier) leaves every cosine within +-0.1 of zero."""
    x = pooled_ref.double()
    x = x / x.norm(dim=1, keepdim=True)
    d = x.mean(0)
    d = d / d.norm()
    dev = x - (x @ d)[:, None] * d[None, :]
    r = torch.linalg.svd(dev, full_matrices=False).Vh[0]          # first principal direction of the deviations
    along = float((x @ d).median())
    a = MARGIN / along
    b = float(np.sqrt(max(1.0 - a * a, 0.0)))
    w1 = a * d + b * r
    w0 = a * d - b * r
    return torch.stack([w0, w1], dim=1).float().numpy() * 0.05    # any column scale: the head normalises it


def _check_margin_parity(got, ref, cos_ref, labels, what, tol=TOL):
    """|got - ref| <= TOL for every pair whose label cosine is not within 2e-3 of the threshold; the others are the
    straddlers SURVEY section 7 asks to flag: a 10.5-logit step sits on them, so they are reported, and may differ."""
    lab = torch.as_tensor(labels).long()
    g = cos_ref[torch.arange(len(lab)), lab]
    straddle = (g - MARGIN).abs() < 2e-3
    above = int((g > MARGIN).sum())
    err = (got - ref).abs().max(dim=1).values
    worst = float(err[~straddle].max()) if (~straddle).any() else 0.0
    print(f"{what}: label cosine range [{float(g.min()):+.3f}, {float(g.max()):+.3f}], {above} above / "
          f"{len(g) - above} below the {MARGIN} margin, {int(straddle.sum())} straddlers "
          f"{[round(float(v), 5) for v in g[straddle]]}; max|dscore| off the threshold = {worst:.2e}")
    assert above > 0 and above < len(g), "the test must exercise both sides of the margin"
    assert worst <= tol
    return straddle


# ---------------------------------------------------------------------------------------------- AM-softmax head
def test_am_softmax_head_margin_branch_operator():
    """mmr_am_softmax_head on constructed pooled rows whose label cosine is exactly placed: -1, 0, 0.3499, 0.3501, 0.9, 1
    for both label values, against model_triple.amsoftmax_loss as restated by the oracle (oracle/imagebert.py)."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code import _runtime
    from oracle import imagebert
    g = torch.Generator().manual_seed(11)
    q, _ = torch.linalg.qr(torch.randn(768, 3, generator=g, dtype=torch.float64))
    w0, w1, u = q[:, 0], q[:, 1], q[:, 2]
    am_kernel = torch.stack([w0 * 0.7, w1 * 1.9], dim=1).float().numpy()      # un-normalised columns, as in a checkpoint
    rows, labels, want_cos = [], [], []
    for lab in (0, 1):
        for c in (-1.0, 0.0, 0.3499, 0.3501, 0.9, 1.0):
            wl = w1 if lab else w0
            other = 0.2 if abs(c) < 0.9 else 0.0                                # some cosine with the other column too
            rest = max(1.0 - c * c - other * other, 0.0) ** 0.5
            x = c * wl + other * (w0 if lab else w1) + rest * u
            rows.append((x * 3.7).float())                                       # any norm: the head normalises
            labels.append(lab)
            want_cos.append(c)
    pooled = torch.stack(rows)
    lab_t = torch.tensor(labels, dtype=torch.int32)
    ref = imagebert.amsoftmax_probs(pooled, lab_t, {"cls/seq_relationship/am_kernel": torch.from_numpy(am_kernel)})
    probs, logits = _runtime.am_softmax_head(pooled.cuda(), am_kernel, lab_t)
    torch.cuda.synchronize()
    probs, logits = probs.cpu(), logits.cpu()
    cos = _cosines(pooled, am_kernel)
    for i, (lab, c) in enumerate(zip(labels, want_cos)):
        assert abs(float(cos[i, lab]) - c) < 1e-6
        took_margin = abs(float(logits[i, lab]) - SCALE * (c - MARGIN)) < 1e-3
        assert took_margin == (c > MARGIN), (lab, c, logits[i])
    assert (probs - ref).abs().max().item() < 2e-6
    assert torch.allclose(probs.sum(1), torch.ones(len(rows)), atol=1e-6)


@pytest.mark.parametrize("precision", ["fast", "strict"])
@pytest.mark.parametrize("depth", ["2-layer", "12-layer"])
def test_zk_margin_branch_full_model(depth, precision):
    """The whole zk model with an am_kernel ALIGNED to the pooled output, so that the label cosines of the batch lie on
    both sides of the 0.35 threshold (with the Xavier initialiser they never leave +-0.1 and the margin branch is dead):
    scores within 1e-3 of the oracle off the threshold, straddlers listed; half of the pairs carry label 0."""
    from oracle import imagebert
    cfg = ModelConfig(ZK, n_layers=2, lq=20, nbox=8, vocab=2000) if depth == "2-layer" else _full_cfg(ZK, vocab=3000)
    B = 48
    w = synth.make_weights(cfg, seed=synth.SEED0 + 21)
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 21, n_queries=4)
    inp["labels"] = (np.arange(B) % 2).astype(np.int32)
    pooled_ref = _oracle(cfg, w, inp)["pooled"]
    w = dict(w)
    w["cls/seq_relationship/am_kernel"] = _aligned_am_kernel(pooled_ref)
    ref = imagebert.amsoftmax_probs(pooled_ref, torch.from_numpy(inp["labels"]), imagebert.to_torch(w))
    sc = _scorer(cfg, w, B, precision=precision)
    try:
        probs, pooled = _gpu(sc, inp)
    finally:
        sc.close()
    cos_ref = _cosines(pooled_ref, w["cls/seq_relationship/am_kernel"])
    cos_gpu = _cosines(pooled, w["cls/seq_relationship/am_kernel"])
    print(f"zk {depth} {precision}: max|dcos| GPU vs oracle = {float((cos_gpu - cos_ref).abs().max()):.2e}; score range "
          f"[{float(ref[:, 1].min()):.4f}, {float(ref[:, 1].max()):.4f}]")
    straddle = _check_margin_parity(probs, ref, cos_ref, inp["labels"], f"zk {depth} {precision}")
    # a straddler either agrees too (same side taken) or differs by the step: nothing in between
    lab = torch.from_numpy(inp["labels"]).long()
    same_side = (cos_gpu[torch.arange(B), lab] > MARGIN) == (cos_ref[torch.arange(B), lab] > MARGIN)
    assert ((probs - ref).abs().max(dim=1).values[straddle & same_side] <= 5 * TOL).all()


# ---------------------------------------------------------------------------------------------- headline sizes
def _assert_topk_tie_aware(ref, got, owner, k=5, tol=TOL):
    """Per query: any two candidates the oracle separates by more than 2 tol keep their order, and the top-k LIST is
    identical whenever the oracle's first k + 1 scores are pairwise separated by more than 2 tol."""
    n_exact = 0
    for q in np.unique(owner):
        idx = np.nonzero(owner == q)[0]
        r, g = ref[idx], got[idx]
        gap = r[:, None] - r[None, :]
        assert not ((gap > 2 * tol) & (g[:, None] <= g[None, :])).any(), f"query {q}: order violated beyond the tolerance"
        order = np.argsort(-r, kind="stable")
        head = r[order[:k + 1]]
        if len(head) > 1 and np.min(-np.diff(head)) > 2 * tol:
            assert (np.argsort(-g, kind="stable")[:k] == order[:k]).all(), f"query {q}: top-{k} differs"
            n_exact += 1
    return n_exact


@pytest.mark.parametrize("kind", [ZK, LDS, LXMERT])
def test_b256_full_depth_oracle_parity(kind):
    """BASELINE configs[1] / [2] themselves: B = 256, 32 x 36 x 2048, 12 layers (9/5/5), full vocabulary -- 68 row blocks
    in three waves of the fused GEMM+LayerNorm kernel and the split tail wave of the 16-bit GEMM -- against the fp32
    oracle (a few seconds of CPU per model)."""
    cfg = _full_cfg(kind)
    B = 256
    w = synth.make_weights(cfg, seed=synth.SEED0 + 3)
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 3, n_queries=9)
    ref = _oracle(cfg, w, inp)
    sc = _scorer(cfg, w, B)
    try:
        probs, pooled = _gpu(sc, inp)
        launches = sc.launches_per_forward()
    finally:
        sc.close()
    err = (probs - ref["probs"]).abs().max().item()
    perr = (pooled - ref["pooled"]).abs().max().item()
    print(f"{kind} B=256 full depth: max|dscore| = {err:.3e}, max|dpooled| = {perr:.3e}, {launches} launches")
    assert err <= TOL
    _assert_topk_tie_aware(ref["probs"][:, 1].numpy(), probs[:, 1].numpy(), inp["query_owner"])


@pytest.mark.parametrize("precision", ["fast", "strict"])
def test_cfg4_shaped_candidate_set_top5(precision):
    """BASELINE configs[3] in miniature: 50 queries x 30 candidates (1,500 pairs, the same query ids repeated over a
    query's candidates), 12-layer zk at 32 x 36, through the chunked host path (five full 256-pair chunks + a ragged
    one).  The am_kernel is aligned to the pooled output so that the scores spread over (0, 1) and cross the margin
    threshold instead of sitting in a 1 %-wide band: scores within 1e-3 off the threshold, tie-aware identical top-5."""
    from oracle import imagebert
    cfg = _full_cfg(ZK, vocab=3000)
    nq, per_q = 50, 30
    N = nq * per_q
    w = synth.make_weights(cfg, seed=synth.SEED0 + 4)
    inp = synth.make_inputs(cfg, N, seed=synth.SEED0 + 4, n_queries=nq)
    wt = imagebert.to_torch(w)
    pooled_ref = torch.cat([_oracle(cfg, w, {k: v[lo:lo + 300] for k, v in inp.items()})["pooled"]
                            for lo in range(0, N, 300)])
    w = dict(w)
    w["cls/seq_relationship/am_kernel"] = _aligned_am_kernel(pooled_ref)
    wt["cls/seq_relationship/am_kernel"] = torch.from_numpy(w["cls/seq_relationship/am_kernel"])
    ref = imagebert.amsoftmax_probs(pooled_ref, torch.from_numpy(inp["labels"]), wt)
    sc = _scorer(cfg, w, 256, precision=precision)
    try:
        got = sc.score(sc.to_feeds(inp)).clone()
    finally:
        sc.close()
    cos_ref = _cosines(pooled_ref, w["cls/seq_relationship/am_kernel"])
    straddle = _check_margin_parity(got, ref, cos_ref, inp["labels"], f"cfg4-shaped zk {precision}")
    r, g = ref[:, 1].numpy().copy(), got[:, 1].numpy().copy()
    keep = ~straddle.numpy()                                   # a straddler may legitimately jump by the margin step
    n_exact = _assert_topk_tie_aware(r[keep], g[keep], inp["query_owner"][keep])
    spread = float(np.percentile(r, 95) - np.percentile(r, 5))
    print(f"cfg4-shaped: score spread (5..95 %) {spread:.3f}; {n_exact} of {nq} queries have an unambiguous top-5 "
          f"(identical on the GPU)")
    assert spread > 0.2 and n_exact > 0


# ---------------------------------------------------------------------------------------------- strict precision
def test_split3_and_fp32_attention_operators():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(5)
    x = torch.randn(37, 256, device="cuda") * torch.logspace(-3, 2, 256, device="cuda")
    for weights in (0, 1):
        out = torch.empty(37, 768, dtype=torch.float16, device="cuda")
        _lib.check(lib.mmr_split3(x.data_ptr(), 256, 37, 256, out.data_ptr(), 768, _lib.ACT_NONE, weights, _lib.DT_FP16, st))
        torch.cuda.synchronize()
        hi = x.half()
        lo = (x - hi.float()).half()
        parts = (hi, hi, lo) if weights else (hi, lo, hi)
        for i, p_ in enumerate(parts):
            assert torch.equal(out[:, 256 * i:256 * (i + 1)], p_)
        # two fp16 terms carry ~21 bits down to fp16's subnormal spacing (6e-8): below ~1e-3 the ABSOLUTE floor rules
        err = (hi.float() + lo.float() - x).abs()
        assert (err <= torch.maximum(x.abs() * 2 ** -20, torch.full_like(x, 6.1e-8))).all(), err.max().item()
    # precise GELU applied while splitting
    out = torch.empty(37, 768, dtype=torch.float16, device="cuda")
    _lib.check(lib.mmr_split3(x.data_ptr(), 256, 37, 256, out.data_ptr(), 768, _lib.ACT_GELU_TANH, 0, _lib.DT_FP16, st))
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(x.double(), approximate="tanh")
    got = out[:, :256].double() + out[:, 256:512].double()
    assert ((got - ref).abs() <= torch.maximum(ref.abs() * 2e-6, torch.full_like(ref, 1.3e-7))).all()

    H = 12
    for (B, Sq, Sk, masked) in [(3, 68, 68, True), (2, 32, 36, True), (2, 36, 32, False), (2, 104, 104, False), (5, 1, 68, True)]:
        q = torch.randn(B * Sq, 3 * 768, device="cuda")
        kv = q if Sq == Sk else torch.randn(B * Sk, 3 * 768, device="cuda")
        mask = None
        if masked:
            lens = torch.randint(1, Sk + 1, (B,), device="cuda")
            mask = (torch.arange(Sk, device="cuda")[None, :] < lens[:, None]).int().contiguous()
        out = torch.empty(B * Sq, 768, device="cuda")
        _lib.check(lib.mmr_attention_f32(q.data_ptr(), Sq * 2304, 2304, kv[:, 768:].data_ptr(), kv[:, 1536:].data_ptr(),
                                         Sk * 2304, 2304, 0 if mask is None else mask.data_ptr(), out.data_ptr(), Sq * 768,
                                         768, B, Sq, Sk, H, st))
        torch.cuda.synchronize()
        qq = q[:, :768].double().view(B, Sq, H, 64).transpose(1, 2)
        kk = kv[:, 768:1536].double().view(B, Sk, H, 64).transpose(1, 2)
        vv = kv[:, 1536:].double().view(B, Sk, H, 64).transpose(1, 2)
        s = qq @ kk.transpose(-1, -2) / 8.0
        if mask is not None:
            s = s + (1.0 - mask.double())[:, None, None, :] * -10000.0
        ref = (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(B * Sq, 768)
        assert (out.double() - ref).abs().max().item() < 2e-5, (B, Sq, Sk)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_cls_attention_operator(dtype):
    """Attention for the first query row of every pair (the last block's [CLS] tail) against fp32 torch and against
    the full tcgen05 attention kernel's first rows."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib, ops
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(6)
    H = 12
    for (B, S, masked) in [(7, 68, True), (3, 32, True), (2, 104, False), (300, 68, True), (2, 1, False)]:
        qkv = torch.randn(B * S, 2304, device="cuda").to(dtype)
        mask = None
        if masked:
            lens = torch.randint(1, S + 1, (B,), device="cuda")
            mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).int().contiguous()
        out = torch.empty(B, 768, dtype=dtype, device="cuda")
        _lib.check(lib.mmr_cls_attention(qkv.data_ptr(), S * 2304, qkv[:, 768:].data_ptr(), qkv[:, 1536:].data_ptr(), 2304,
                                         0 if mask is None else mask.data_ptr(), out.data_ptr(), 768, B, S, H,
                                         _lib.DT_BF16 if dtype == torch.bfloat16 else _lib.DT_FP16, st))
        torch.cuda.synchronize()
        q = qkv[::S, :768].float().view(B, 1, H, 64).transpose(1, 2)
        k = qkv[:, 768:1536].float().view(B, S, H, 64).transpose(1, 2)
        v = qkv[:, 1536:].float().view(B, S, H, 64).transpose(1, 2)
        s = q @ k.transpose(-1, -2) / 8.0
        if mask is not None:
            s = s + (1.0 - mask.float())[:, None, None, :] * -10000.0
        ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, 768)
        rel = ((out.float() - ref).abs().max() / ref.abs().max()).item()
        assert rel < (5e-3 if dtype == torch.bfloat16 else 1.5e-3), (B, S, rel)
        full = ops.attention(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], mask, B, S, S, H)
        torch.cuda.synchronize()
        d = (out.float() - full[::S].float()).abs().max().item()
        assert d <= (8e-3 if dtype == torch.bfloat16 else 1e-3) * ref.abs().max().item() + 1e-6, (B, S, d)


STRICT_CASES = [("cfg1", ZK), ("cfg1", LDS), ("cfg1", LXMERT), ("golden", "zk"), ("golden", "lds"), ("golden", "lxmert"),
                ("full", ZK), ("full", LDS), ("full", LXMERT)]


@pytest.mark.parametrize("case,kind", STRICT_CASES)
def test_strict_precision_meets_1e3_on_trained_like_weights(case, kind):
    """The weight set on which 16-bit operands miss the stated tolerance ("trained-like": every matrix x3, random
    LayerNorm affine; 1.3e-3..2.9e-3 in fast mode) under precision="strict": |dscore| <= 1e-3 -- in fact two orders
    below it -- for the 2-layer plumbing config, the reference-code goldens and the 12-layer / 9-5-5 models at 32 x 36."""
    if case == "golden":
        g = np.load(os.path.join(GOLD, f"{kind}_ref_shim_small_trained.npz" if kind != "lxmert"
                                 else "lxmert_ref_small_trained.npz"))
        cfg = ModelConfig(**ast.literal_eval(str(g["cfg"])))
        B = int(g["batch"])
        w = synth.make_weights(cfg, seed=int(g["seed"]), trained_like=True)
        inp = synth.make_inputs(cfg, B, seed=int(g["seed"]))
        ref = torch.from_numpy(g["probs"])
    else:
        if case == "cfg1":
            cfg = (ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=2, lq=20, nbox=8, vocab=2000) if kind == LXMERT
                   else ModelConfig(kind, n_layers=2, lq=20, nbox=8, vocab=2000))
            B = 4
        else:
            cfg, B = _full_cfg(kind, vocab=3000), 24
        w = synth.make_weights(cfg, seed=synth.SEED0 + 1, trained_like=True)
        inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 1, n_queries=2)
        ref = _oracle(cfg, w, inp)["probs"]
    errs = {}
    for precision in ("fast", "strict"):
        sc = _scorer(cfg, w, B, precision=precision)
        try:
            probs, _ = _gpu(sc, inp)
            errs[precision] = (probs - ref).abs().max().item()
        finally:
            sc.close()
    print(f"{case} {kind} trained-like: max|dscore| fast = {errs['fast']:.3e}, strict = {errs['strict']:.3e}")
    assert errs["strict"] <= TOL
    assert errs["strict"] <= 1e-4          # what the two-term split actually achieves, with margin
    assert errs["fast"] <= 4e-3            # the fast path's documented bound on this weight set


@pytest.mark.parametrize("kind", [ZK, LXMERT])
def test_strict_precision_prunes_and_taps_like_the_fast_path(kind):
    """strict mode through the same switches: full last block == [CLS]-only last block, taps available."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib
    lib = _lib.load()
    cfg = (ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=2, lq=20, nbox=8, vocab=2000) if kind == LXMERT
           else ModelConfig(kind, n_layers=2, lq=20, nbox=8, vocab=2000))
    B = 6
    w = synth.make_weights(cfg, seed=synth.SEED0 + 31, trained_like=True)
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 31, n_queries=2)
    ref = _oracle(cfg, w, inp)
    sc = _scorer(cfg, w, B, precision="strict")
    try:
        out = {}
        for prune in (1, 0):
            _lib.check(lib.mmr_set_tuning(_lib.TUNE_PRUNE_LAST, prune))
            out[prune], _ = _gpu(sc, inp)
        assert (out[0] - out[1]).abs().max().item() <= 2e-6
        sc.set_debug_taps(1)
        _gpu(sc, inp)
        seq = sc.activation(1, B).cpu()
        H = cfg.hidden
        ref_seq = (torch.cat([ref["lang"].reshape(-1, H), ref["visn"].reshape(-1, H)]) if kind == LXMERT
                   else ref["sequence_output"].reshape(-1, H))
        assert (seq - ref_seq).abs().max().item() < 2e-3      # every row of the last block, near-fp32
    finally:
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_PRUNE_LAST, 1))
        sc.close()


# ---------------------------------------------------------------------------------------------- pruned last block
@pytest.mark.parametrize("kind,B", [(ZK, 24), (LDS, 24), (LXMERT, 24), (ZK, 256), (LXMERT, 256)])
def test_last_block_for_cls_rows_only_matches_the_full_block(kind, B):
    """MMR_TUNE_PRUNE_LAST: keys / values of the last block for all rows, attention + output projection + FFN for the
    [CLS] rows only (LXMERT: the last cross layer's visual half dropped); final-layer tap refused unless taps are on.
    The two paths are the same arithmetic in a different accumulation ORDER (one-row attention on the CUDA cores, row
    LayerNorm kernel instead of the fused epilogue): fp32-level differences in the logits flip the 16-bit rounding of
    about one attention probability per pair, which moves a head's context row by < 1 ulp and re-rounds ~40 % of it --
    measured 4e-5..1.4e-4 on the score, the size of the fp16 path's own distance to the fp32 oracle (5e-4), and 2e-6 in
    strict mode (test_strict_precision_prunes_and_taps_like_the_fast_path).  Both must sit inside the tolerance."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib
    lib = _lib.load()
    cfg = _full_cfg(kind, vocab=3000)
    w = synth.make_weights(cfg, seed=synth.SEED0 + 17)
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 17, n_queries=3)
    sc = _scorer(cfg, w, B)
    try:
        out, pooled, launches = {}, {}, {}
        for prune in (1, 0):
            _lib.check(lib.mmr_set_tuning(_lib.TUNE_PRUNE_LAST, prune))
            out[prune], pooled[prune] = _gpu(sc, inp)
            launches[prune] = sc.launches_per_forward()
        d = (out[0] - out[1]).abs().max().item()
        dp = (pooled[0] - pooled[1]).abs().max().item()
        print(f"{kind} B={B}: pruned vs full last block: max|dscore| = {d:.2e}, max|dpooled| = {dp:.2e}; launches "
              f"{launches[1]} (pruned) vs {launches[0]} (full)")
        assert d <= 4e-4
        if B <= 24:
            ref = _oracle(cfg, w, inp)["probs"]
            assert (out[1] - ref).abs().max().item() <= TOL and (out[0] - ref).abs().max().item() <= TOL
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_PRUNE_LAST, 1))
        _gpu(sc, inp)
        with pytest.raises(_lib.MmrError, match="final-layer tap"):
            sc.activation(1, B)
    finally:
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_PRUNE_LAST, 1))
        sc.close()


# ---------------------------------------------------------------------------------------------- several handles
def test_handles_do_not_share_exchange_tables_or_graphs():
    """the fused GEMM+LayerNorm exchange table used to be one per device, reallocated when a larger
    handle appeared -- captured graphs of the smaller handle then replayed onto freed memory.  Now every handle owns its
    table: small scorer (graphs captured), then larger ones of other kinds, then the small one again, bit-identical."""
    cfgs = [ModelConfig(ZK, n_layers=2, lq=32, nbox=36, vocab=2000), ModelConfig(LDS, n_layers=2, lq=32, nbox=36, vocab=2000)]
    B_small, B_big = 8, 40
    w = [synth.make_weights(c, seed=synth.SEED0 + 41 + i) for i, c in enumerate(cfgs)]
    small = _scorer(cfgs[0], w[0], B_small)
    feeds = {k: v.cuda() for k, v in small.to_feeds(synth.make_inputs(cfgs[0], B_small, seed=9)).items()}
    out = torch.empty((B_small, 2), dtype=torch.float32, device="cuda")
    first = small.forward_device(feeds, probs_out=out).clone()          # eager
    small.forward_device(feeds, probs_out=out)                          # captured
    big = [_scorer(c, ww, B_big) for c, ww in zip(cfgs, w)]
    try:
        for sc in big:
            inp = synth.make_inputs(sc.cfg, B_big, seed=10)
            sc.forward_device({k: v.cuda() for k, v in sc.to_feeds(inp).items()})
        for _ in range(3):
            out.zero_()
            small.forward_device(feeds, probs_out=out)                  # replayed
            torch.cuda.synchronize()
            assert torch.equal(out, first)
        # two handles on two streams at once: the library serialises them on the device; results as when run alone
        alone = []
        inps = [{k: v.cuda() for k, v in sc.to_feeds(synth.make_inputs(sc.cfg, B_big, seed=12 + i)).items()}
                for i, sc in enumerate(big)]
        for sc, f in zip(big, inps):
            alone.append(sc.forward_device(f).clone())
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream() for _ in big]
        both = [None, None]
        for rep in range(4):
            for i, (sc, f, s) in enumerate(zip(big, inps, streams)):
                with torch.cuda.stream(s):
                    both[i] = sc.forward_device(f).clone()
        torch.cuda.synchronize()
        for a, b in zip(alone, both):
            assert torch.equal(a, b)
    finally:
        small.close()
        for sc in big:
            sc.close()


def test_cast16_saturates_instead_of_overflowing():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    x = torch.tensor([1e6, -1e6, 65504.0, 70000.0, 3.0, -0.0, 1e-9, 65519.0] * 4, device="cuda")
    y = ops.cast16(x, torch.float16)
    assert torch.isfinite(y).all()
    assert y[0].item() == 65504.0 and y[1].item() == -65504.0 and y[3].item() == 65504.0 and y[4].item() == 3.0
    assert torch.isinf(ops.cast16(x, torch.bfloat16)).sum() == 0      # bf16 has fp32's range


# ---------------------------------------------------------------------------------------------- two problems, one launch
def test_lxmert_paired_attention_launch_is_bit_identical():
    """LXMERT issues attention in pairs -- both streams' self-attention, both directions of the shared cross-attention
    block -- and the tcgen05 kernel takes the two problems as two SEGMENTS of one launch (different Sq / Sk / mask /
    output).  Same bits as two launches (the arithmetic of an item does not depend on what else the launch carries),
    fewer launches, within tolerance of the oracle; 32 x 36 (3 key chunks for both) and 20 x 40 (2 vs 3 chunks: the
    shorter segment's stage rows hold the longer one's stale keys behind a -inf mask)."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib
    lib = _lib.load()
    # (16 x 12: one key chunk for both segments -> the generic kernel; 192 visual rows keep the visual stream on the fused
    # GEMM+LayerNorm path in both modes, which needs more than 128 rows -- the comparison is about attention)
    for lq, nbox, B in ((32, 36, 16), (20, 40, 13), (16, 12, 16)):
        cfg = ModelConfig(LXMERT, n_layers=3, n_r_layers=2, n_x_layers=2, lq=lq, nbox=nbox, vocab=2000)
        w = synth.make_weights(cfg, seed=synth.SEED0 + 61)
        inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 61, n_queries=2)
        sc = _scorer(cfg, w, B)
        try:
            out, launches = {}, {}
            _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_QUERY_DEDUP, 0))     # the same route for the language stream in both modes
            for merged in (1, 0):
                _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_MERGE, merged))
                _lib.check(lib.mmr_set_tuning(_lib.TUNE_PRUNE_LAST, 0))     # every attention pair of the graph
                out[merged], _ = _gpu(sc, inp)
                launches[merged] = sc.launches_per_forward()
            assert torch.equal(out[0], out[1]), (lq, nbox)
            assert launches[1] < launches[0]
            assert (out[1] - _oracle(cfg, w, inp)["probs"]).abs().max().item() <= TOL
        finally:
            _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_MERGE, 1))
            _lib.check(lib.mmr_set_tuning(_lib.TUNE_PRUNE_LAST, 1))
            _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_QUERY_DEDUP, 1))
            sc.close()


# ---------------------------------------------------------------------------------------------- language stream once per query
@pytest.mark.parametrize("B,nq,full", [(60, 2, False), (256, 9, True), (17, 17, False), (33, 1, False)])
def test_lxmert_language_blocks_once_per_distinct_query(B, nq, full):
    """LXMERT's first n_layers blocks depend on the query only (modeling.py:577-578): with mmr_inputs.lang_unique /
    lang_slot they run once per distinct query of the batch and are expanded before the cross-modality blocks.  Same
    scores as every pair computing its own (bit-identical where both paths take the same kernels: more than 128 compact
    rows), within tolerance of the oracle; a batch of all-distinct queries takes the ordinary path; the host path
    (score: chunks of max_batch, grouping per chunk) agrees with the device path."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import distinct_queries
    lib = _lib.load()
    cfg = _full_cfg(LXMERT, vocab=3000) if full else ModelConfig(LXMERT, n_layers=3, n_r_layers=2, n_x_layers=2, lq=32,
                                                                  nbox=36, vocab=2000)
    w = synth.make_weights(cfg, seed=synth.SEED0 + 71)
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 71, n_queries=nq)
    uniq, slot = distinct_queries(inp["query_ids"], inp["query_mask"])
    assert len(uniq) <= nq and (inp["query_ids"][uniq.numpy()[slot.numpy()]] == inp["query_ids"]).all()
    sc = _scorer(cfg, w, B)
    try:
        out, launches = {}, {}
        for dedup in (1, 0):
            _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_QUERY_DEDUP, dedup))
            out[dedup], _ = _gpu(sc, inp)
            launches[dedup] = sc.launches_per_forward()
        d = (out[0] - out[1]).abs().max().item()
        print(f"lxmert B={B}, {len(uniq)} distinct queries: dedup vs per-pair max|dscore| = {d:.2e}; launches "
              f"{launches[1]} vs {launches[0]}")
        if len(uniq) < B:      # the compact stream rides the merged launches (padded to whole tiles): same kernels, same bits
            assert torch.equal(out[0], out[1])
        assert d <= 4e-4
        if not full:
            assert (out[1] - _oracle(cfg, w, inp)["probs"]).abs().max().item() <= TOL
        host = sc.to_feeds(inp)
        small = _scorer(cfg, w, 16)
        try:
            got = small.score(host)                               # ragged chunks, grouping recomputed per chunk
        finally:
            small.close()
        assert (got - out[1]).abs().max().item() <= 4e-4
    finally:
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_QUERY_DEDUP, 1))
        sc.close()


@pytest.mark.parametrize("kind", [ZK, LDS, LXMERT])
def test_bf16_operands_in_strict_mode_meet_the_tolerance(kind):
    """bf16 operands (the north star's operand type) miss 1e-3 on the zk head when rounded once (5e-3..1.5e-2,
    DESIGN.md section 2) -- which is why the default operand type is fp16; with two-term split operands (8 + 8
    significand bits) they meet it on the 12-layer / 9-5-5 models at 32 x 36."""
    cfg, B = _full_cfg(kind, vocab=3000), 24
    w = synth.make_weights(cfg, seed=synth.SEED0 + 1)
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 1, n_queries=2)
    ref = _oracle(cfg, w, inp)["probs"]
    errs = {}
    for precision in ("fast", "strict"):
        sc = _scorer(cfg, w, B, dtype="bf16", precision=precision)
        try:
            errs[precision] = (_gpu(sc, inp)[0] - ref).abs().max().item()
        finally:
            sc.close()
    print(f"{kind} bf16 operands: max|dscore| fast = {errs['fast']:.3e}, strict = {errs['strict']:.3e}")
    assert errs["strict"] <= TOL
    assert errs["fast"] <= 3e-2


@pytest.mark.parametrize("kind", [ZK, LDS, LXMERT])
def test_odd_batch_sizes_through_the_cls_tail(kind):
    """B = 1, 7 and 9 pairs (not multiples of the 8-row clusters of the fused pooler + head kernel, a single pair, a
    max_batch larger than the batch) against the oracle, native 20 x 10 shapes, 2 layers."""
    cfg = (ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=2, lq=23, nbox=10, vocab=2000) if kind == LXMERT
           else ModelConfig(kind, n_layers=2, lq=20, nbox=10, vocab=2000))
    w = synth.make_weights(cfg, seed=synth.SEED0 + 81)
    sc = _scorer(cfg, w, 16)
    try:
        for B in (1, 7, 9):
            inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 81 + B, n_queries=max(1, B // 3))
            ref = _oracle(cfg, w, inp)
            probs, pooled = _gpu(sc, inp)
            assert (probs - ref["probs"]).abs().max().item() <= TOL, (kind, B)
            assert (pooled - ref["pooled"]).abs().max().item() <= 5e-3, (kind, B)
            assert torch.allclose(probs.sum(1), torch.ones(B), atol=1e-6)
    finally:
        sc.close()
