"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the fp32 CPU
oracle on the same seeded inputs, against the committed golden vectors of the reference's own LXMERT code, and
through size-independent properties at the BASELINE sizes.

Stated tolerance (BASELINE.json north_star): |score - oracle| <= 1e-3 absolute, identical top-k ordering.
Operands are fp16 (fp32 accumulate / residual / LayerNorm / softmax): bf16 operands measure 5e-3..1.5e-2 on the zk
AM-softmax head (profiles/r01_numerics_budget_cpu_emulation.log) and cannot meet 1e-3.
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
TOL = 1e-3          # the stated tolerance: weights from the reference's own initialisers
TOL_STRESS = 4e-3   # "trained-like" stress weights (every matrix x3, random LN affine): they amplify ANY operand
                    # rounding (the CPU emulation of fp16 operands gives the same 1.3e-3..2.9e-3), see DESIGN.md section 2


def _scorer(cfg, w, max_batch, dtype="fp16"):
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import MatchScorer
    return MatchScorer(cfg, w, device=0, dtype=dtype, max_batch=max_batch)


def _oracle(cfg, w, inp):
    from oracle import imagebert, lxmert
    wt, it = imagebert.to_torch(w), imagebert.to_torch(inp)
    if cfg.kind == ZK:
        return imagebert.zk_forward(wt, it, cfg.n_layers)
    if cfg.kind == LDS:
        return imagebert.lds_forward(wt, it, cfg.n_layers)
    return lxmert.forward(wt, it, cfg.n_layers, cfg.n_r_layers, cfg.n_x_layers)


def _gpu_probs(sc, inp, taps=False):
    feeds = {k: v.cuda() for k, v in sc.to_feeds(inp).items()}
    if taps:
        sc.set_debug_taps(True)
    pooled = torch.empty((feeds["query_ids"].shape[0], sc.cfg.hidden), dtype=torch.float32, device="cuda")
    probs = sc.forward_device(feeds, pooled_out=pooled)
    torch.cuda.synchronize()
    return probs.cpu(), pooled.cpu()


def _small_cfg(kind, **kw):
    if kind == LXMERT:
        return ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=2, lq=20, nbox=8, vocab=2000, **kw)
    return ModelConfig(kind, n_layers=2, lq=20, nbox=8, vocab=2000, **kw)


@pytest.mark.parametrize("trained_like", [False, True])
@pytest.mark.parametrize("kind", [ZK, LDS, LXMERT])
def test_cfg1_parity_with_activation_taps(kind, trained_like):
    """BASELINE configs[0]: 1 query x 8 regions x 2048-d, 2-layer, batch 4 (one query scored against 4 products)."""
    cfg = _small_cfg(kind)
    w = synth.make_weights(cfg, seed=synth.SEED0, trained_like=trained_like)
    inp = synth.make_inputs(cfg, 4, seed=synth.SEED0, n_queries=1)
    ref = _oracle(cfg, w, inp)
    sc = _scorer(cfg, w, 4)
    probs, pooled = _gpu_probs(sc, inp, taps=True)
    emb = sc.activation(0, 4).cpu()
    seq = sc.activation(1, 4).cpu()
    H = cfg.hidden
    if kind == LXMERT:
        ref_emb = torch.cat([ref["embedding_output"].reshape(-1, H), ref["visn_embedding"].reshape(-1, H)])
        ref_seq = torch.cat([ref["lang"].reshape(-1, H), ref["visn"].reshape(-1, H)])
    else:
        ref_emb = ref["embedding_output"].reshape(-1, H)
        ref_seq = ref["sequence_output"].reshape(-1, H)
    # embeddings: fp32 gathers + LayerNorm, plus one 16-bit-operand projection for the region rows
    assert (emb - ref_emb).abs().max().item() < (2e-2 if trained_like else 5e-3)
    assert (seq - ref_seq).abs().max().item() < (5e-2 if trained_like else 1e-2)
    assert (pooled - ref["pooled"]).abs().max().item() < 2e-2
    err = (probs - ref["probs"]).abs().max().item()
    print(f"{kind} trained_like={trained_like}: max|dscore| = {err:.3e}")
    assert err <= (TOL_STRESS if trained_like else TOL)
    assert torch.allclose(probs.sum(1), torch.ones(4), atol=1e-6)


@pytest.mark.parametrize("tag", ["small", "small_trained", "native", "cfg3shape"])
def test_lxmert_against_reference_golden(tag):
    """CUDA LXMERT vs outputs of the reference's OWN KDDModel.forward (tests/golden, tools/make_golden.py)."""
    g = np.load(os.path.join(GOLD, f"lxmert_ref_{tag}.npz"))
    cfg = ModelConfig(**ast.literal_eval(str(g["cfg"])))
    B = int(g["batch"])
    w = synth.make_weights(cfg, seed=int(g["seed"]), trained_like=bool(g["trained_like"]))
    inp = synth.make_inputs(cfg, B, seed=int(g["seed"]))
    sc = _scorer(cfg, w, B)
    probs, pooled = _gpu_probs(sc, inp)
    err = np.abs(probs.numpy() - g["probs"]).max()
    print(f"lxmert golden {tag}: max|dscore| = {err:.3e}")
    assert err <= (TOL_STRESS if bool(g["trained_like"]) else TOL)
    xn = pooled / pooled.norm(dim=1, keepdim=True)
    assert np.abs(xn.numpy() - g["x_norm"]).max() < (6e-3 if bool(g["trained_like"]) else 2e-3)


@pytest.mark.parametrize("tag", ["small", "small_trained", "native"])
@pytest.mark.parametrize("kind", ["zk", "lds"])
def test_imagebert_against_reference_code_golden(kind, tag):
    """CUDA zk / lds vs outputs of the reference's OWN TF-1 model code executed on the eager TensorFlow-op stand-in
    (tests/golden/{zk,lds}_ref_shim_*.npz, tools/make_golden.py --tf-shim): native 20 x 10 shapes, 2 and 12 layers."""
    g = np.load(os.path.join(GOLD, f"{kind}_ref_shim_{tag}.npz"))
    cfg = ModelConfig(**ast.literal_eval(str(g["cfg"])))
    B = int(g["batch"])
    w = synth.make_weights(cfg, seed=int(g["seed"]), trained_like=bool(g["trained_like"]))
    inp = synth.make_inputs(cfg, B, seed=int(g["seed"]))
    sc = _scorer(cfg, w, B)
    try:
        probs, pooled = _gpu_probs(sc, inp)
    finally:
        sc.close()
    err = np.abs(probs.numpy() - g["probs"]).max()
    perr = np.abs(pooled.numpy() - g["pooled"]).max()
    print(f"{kind} reference-code golden {tag}: max|dscore| = {err:.3e}, max|dpooled| = {perr:.3e}")
    assert err <= (TOL_STRESS if bool(g["trained_like"]) else TOL)
    assert perr <= (3e-2 if bool(g["trained_like"]) else 5e-3)


@pytest.mark.parametrize("kind", [ZK, LDS, LXMERT])
def test_full_depth_parity_and_topk_order(kind):
    """12-layer (9/5/5 for LXMERT) at the BASELINE shapes 32 x 36 x 2048: 2 queries x 12 candidates; scores within
    1e-3 of the oracle and identical per-query ranking (pairs closer than 2e-3 in the oracle are not order-checked:
    the tolerance itself allows them to swap)."""
    if kind == LXMERT:
        cfg = ModelConfig(LXMERT, n_layers=9, n_r_layers=5, n_x_layers=5, lq=32, nbox=36, vocab=3000)
    else:
        cfg = ModelConfig(kind, n_layers=12, lq=32, nbox=36, vocab=3000)
    B, nq = 24, 2
    w = synth.make_weights(cfg, seed=synth.SEED0 + 1)
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 1, n_queries=nq)
    ref = _oracle(cfg, w, inp)["probs"][:, 1]
    sc = _scorer(cfg, w, B)
    probs, _ = _gpu_probs(sc, inp)
    got = probs[:, 1]
    err = (got - ref).abs().max().item()
    print(f"{kind}: max|dscore| = {err:.3e}")
    assert err <= TOL
    owner = inp["query_owner"]
    for q in range(nq):
        idx = np.nonzero(owner == q)[0]
        r, g_ = ref[idx].numpy(), got[idx].numpy()
        order_ref = np.argsort(-r, kind="stable")
        gaps = np.abs(np.diff(r[order_ref]))
        if gaps.min() > 2 * TOL:
            assert (np.argsort(-g_, kind="stable") == order_ref).all()
        else:  # ordering must still agree wherever the oracle separates two items by more than the tolerance
            for a in range(len(idx)):
                for b in range(len(idx)):
                    if r[a] - r[b] > 2 * TOL:
                        assert g_[a] > g_[b]


@pytest.mark.parametrize("kind", [ZK, LDS, LXMERT])
def test_baseline_size_properties(kind):
    """BASELINE configs[1]/[2] sizes (B=256, 32 x 36 x 2048, full depth), no oracle: batch-permutation invariance,
    chunking invariance (256 = 2 x 128 gives bit-identical scores: no cross-pair op exists), zk padding invariance,
    probabilities sum to one."""
    if kind == LXMERT:
        cfg = ModelConfig(LXMERT, n_layers=9, n_r_layers=5, n_x_layers=5, lq=32, nbox=36)
    else:
        cfg = ModelConfig(kind, n_layers=12, lq=32, nbox=36)
    B = 256
    w = synth.make_weights(cfg, seed=synth.SEED0 + 2)
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 2)
    sc = _scorer(cfg, w, B)
    feeds = {k: v.cuda() for k, v in sc.to_feeds(inp).items() if k in sc.spec}   # (no per-batch query grouping: sliced below)
    full = sc.forward_device(feeds).clone()
    halves = torch.cat([sc.forward_device({k: v[:128] for k, v in feeds.items()}).clone(),
                        sc.forward_device({k: v[128:] for k, v in feeds.items()}).clone()])
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).cuda()
    permuted = sc.forward_device({k: v[perm].contiguous() for k, v in feeds.items()}).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(full).all()
    assert torch.allclose(full.sum(1), torch.ones(B, device="cuda"), atol=1e-5)
    assert torch.equal(full, halves)
    assert torch.equal(full[perm], permuted)
    assert full[:, 1].std().item() > 0  # scores actually depend on the inputs
    if kind == ZK:
        f2 = feeds["feats"].clone()
        nb = feeds["num_boxes"]
        pad = torch.arange(cfg.nbox, device="cuda")[None, :] >= nb[:, None]
        f2[pad] += 1.0
        out2 = sc.forward_device({**feeds, "feats": f2})
        torch.cuda.synchronize()
        # padded boxes are masked as attention KEYS; the [CLS] row never sees them
        assert (out2 - full).abs().max().item() < 1e-5


def test_host_buffer_path_matches_device_path():
    """MatchScorer.score (pinned host feeds, double-buffered H2D, chunks of max_batch, ragged tail) == forward_device."""
    cfg = _small_cfg(ZK)
    w = synth.make_weights(cfg, seed=5)
    inp = synth.make_inputs(cfg, 37, seed=5)
    sc = _scorer(cfg, w, 8)
    host = sc.to_feeds(inp)
    got = sc.score(host)
    feeds = {k: v.cuda() for k, v in host.items()}
    want = torch.cat([sc.forward_device({k: v[i:i + 8] for k, v in feeds.items()}).cpu() for i in range(0, 37, 8)])
    assert torch.equal(got, want)


def test_bf16_operands_also_run():
    """bf16 is kept as an operand type (same tensor-core rate); its tolerance is looser (documented in DESIGN.md)."""
    cfg = _small_cfg(LDS)
    w = synth.make_weights(cfg, seed=7)
    inp = synth.make_inputs(cfg, 4, seed=7)
    ref = _oracle(cfg, w, inp)["probs"]
    sc = _scorer(cfg, w, 4, dtype="bf16")
    probs, _ = _gpu_probs(sc, inp)
    assert (probs - ref).abs().max().item() < 1e-2


def test_error_behaviour():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200._lib import MmrError
    cfg = _small_cfg(LDS)
    w = synth.make_weights(cfg, seed=3)
    sc = _scorer(cfg, w, 4)
    inp = synth.make_inputs(cfg, 6, seed=3)
    feeds = {k: v.cuda() for k, v in sc.to_feeds(inp).items()}
    with pytest.raises(MmrError, match="max_batch"):
        sc.forward_device(feeds)
    with pytest.raises(ValueError, match="feed 'feats'"):
        sc.forward_device({**{k: v[:4] for k, v in feeds.items()}, "feats": feeds["feats"][:4, :, :100]})
    bad = dict(w)
    del bad["bert/pooler/dense/kernel"]
    with pytest.raises(MmrError, match="bert/pooler/dense/kernel"):
        _scorer(cfg, bad, 4)
    bad = dict(w)
    bad["featureemb/fully_connected/biases"] = np.zeros(5, np.float32)
    with pytest.raises(MmrError, match="featureemb/fully_connected/biases"):
        _scorer(cfg, bad, 4)


def test_cfg5_three_model_ensemble_topk_against_oracle():
    """BASELINE configs[4] at test size: the same candidate pairs scored by imagebert_zk (twice: the reference feeds the
    plain and the sen2forest-rewritten query file through the same model), imagebert_lds and lxmert, merged by the
    main.py ensemble; merged scores within the tolerance of the oracle's, and identical top-5 lists wherever the oracle
    separates neighbours by more than the tolerance allows to swap."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ensemble
    from oracle import ensemble as oracle_ensemble
    nq, per_q = 6, 8
    B = nq * per_q
    shapes = dict(lq=20, nbox=10, vocab=2000)
    cfgs = {ZK: ModelConfig(ZK, n_layers=2, **shapes), LDS: ModelConfig(LDS, n_layers=2, **shapes),
            LXMERT: ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=1, **shapes)}
    got, ref = {}, {}
    for kind, cfg in cfgs.items():
        w = synth.make_weights(cfg, seed=synth.SEED0 + 5, trained_like=True)   # spread-out scores
        inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 5, n_queries=nq)
        sc = _scorer(cfg, w, B)
        got[kind] = _gpu_probs(sc, inp)[0][:, 1].numpy().astype(np.float64)
        ref[kind] = _oracle(cfg, w, inp)["probs"][:, 1].numpy().astype(np.float64)
        sc.close()
    qids = [f"q{i // per_q}" for i in range(B)]
    pids = [f"p{i}" for i in range(B)]                       # every pair its own product: the uniqueness filter passes

    def merged(scores):
        d = {k: ensemble.scores_from_arrays(qids, pids, v) for k, v in scores.items()}
        return ensemble.merge_and_select(d[ZK], d[ZK], d[LDS], d[LXMERT])

    rows_g, merged_g = merged(got)
    o = {k: oracle_ensemble.merge_and_select.__globals__["OrderedDict"]() for k in ref}
    for k, v in ref.items():
        for q, p_, s in zip(qids, pids, v):
            o[k].setdefault(q, {})[p_] = float(s)
    rows_o, merged_o = oracle_ensemble.merge_and_select(o[ZK], {q: dict(r) for q, r in o[ZK].items()}, o[LDS], o[LXMERT])
    tol = TOL_STRESS                                          # trained-like weights (see the header of this file)
    assert merged_g.shape == (B,)                             # flat, in the (query, candidate) order of the inputs
    for i, (q, p_) in enumerate(zip(qids, pids)):
        assert abs(merged_g[i] - merged_o[q][p_]) <= tol
    top_o, top_g = dict(rows_o), dict(rows_g)
    assert set(top_o) == set(top_g) and len(top_o) == nq
    for q in top_o:
        s = sorted(merged_o[q].values(), reverse=True)
        if min(a - b for a, b in zip(s[:6], s[1:7])) > 2 * tol:      # unambiguous at the stated tolerance
            assert top_g[q] == top_o[q], q
        else:
            assert len(set(top_g[q]) & set(top_o[q])) >= 4


@pytest.mark.parametrize("phrases", ["pool", "all_distinct", "one"])
def test_zk_label_term_once_per_phrase_is_bit_identical(phrases):
    """The zk label-text term evaluated once per distinct label phrase of the batch (hash claim + representative +
    per-box sum) gives the same bits as the per-box evaluation: few phrases (the synthetic pool), every box its own
    phrase (no sharing, worst case for the table), and one phrase for all boxes (every box races for one slot)."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib
    lib = _lib.load()
    cfg = ModelConfig(ZK, n_layers=1, lq=20, nbox=10, vocab=2000)
    w = synth.make_weights(cfg, seed=synth.SEED0 + 9)
    B = 96
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 9, n_queries=4)
    rng = np.random.default_rng(7)
    ids = np.array(inp["label_ids"], copy=True)
    if phrases == "all_distinct":
        ids = rng.integers(1, cfg.vocab, size=ids.shape).astype(ids.dtype)
    elif phrases == "one":
        ids[...] = ids.reshape(-1, 8)[0]
    inp = dict(inp, label_ids=ids)
    sc = _scorer(cfg, w, B)
    try:
        out = {}
        for dedup in (1, 0):
            _lib.check(lib.mmr_set_tuning(_lib.TUNE_LABEL_DEDUP, dedup))
            for rep in range(2):   # twice: the phrase table is reused across forwards (epoch tags)
                out[dedup], _ = _gpu_probs(sc, inp)
        assert torch.equal(out[0], out[1])
        assert (out[1] - _oracle(cfg, w, inp)["probs"]).abs().max().item() <= TOL
    finally:
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_LABEL_DEDUP, 1))
        sc.close()


@pytest.mark.parametrize("B,lq", [(8, 32), (16, 16), (12, 20)])
def test_lxmert_two_stream_merged_launches(B, lq):
    """LXMERT's language and visual streams share the activation buffers but not the weights; when batch x query
    length is a multiple of the 256-row tile their projections run as ONE launch each (two weight matrices over a
    row split).  Same bits as the per-stream launches, within tolerance of the oracle; B x lq = 240 exercises the
    fall-back (split not tile-aligned).  Unequal layer counts: 3 language / 2 visual layers pair up twice."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib
    lib = _lib.load()
    cfg = ModelConfig(LXMERT, n_layers=3, n_r_layers=2, n_x_layers=2, lq=lq, nbox=36, vocab=2000)
    w = synth.make_weights(cfg, seed=synth.SEED0 + 11)
    inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 11, n_queries=2)
    sc = _scorer(cfg, w, B)
    try:
        out, launches = {}, {}
        # (query grouping off: with it the language stream takes a different route in the two modes -- riding the merged
        # launches vs its own small ones -- and this test is about the launches themselves)
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_QUERY_DEDUP, 0))
        for merged in (1, 0):
            _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_MERGE, merged))
            out[merged], _ = _gpu_probs(sc, inp)
            launches[merged] = sc.launches_per_forward()
        assert torch.equal(out[0], out[1])
        # fewer launches either way: the GEMM merge needs a tile-aligned split, the paired cross-attention launch does not
        assert launches[1] < launches[0]
        assert (out[1] - _oracle(cfg, w, inp)["probs"]).abs().max().item() <= TOL
    finally:
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_MERGE, 1))
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_LX_QUERY_DEDUP, 1))
        sc.close()


@pytest.mark.parametrize("kind", [ZK, LXMERT])
def test_forward_replays_from_a_cuda_graph(kind):
    """mmr_forward neither allocates nor synchronises, and the state it keeps between forwards (the LayerNorm exchange
    epoch, the label-phrase table epoch) lives on the device: a forward captured into a CUDA graph replays on new
    inputs (different label phrases included) with the bits of an eager forward."""
    cfg = _small_cfg(kind)
    w = synth.make_weights(cfg, seed=synth.SEED0 + 13)
    B = 16
    sc = _scorer(cfg, w, B)
    try:
        sets = []
        for i in range(3):
            inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 100 + i, n_queries=2)
            if i == 2:
                ids = np.array(inp["label_ids"], copy=True)
                ids[...] = np.random.default_rng(i).integers(1, cfg.vocab, size=ids.shape)
                inp = dict(inp, label_ids=ids.astype(np.int32))
            sets.append({k: v.cuda() for k, v in sc.to_feeds(inp).items()})
        eager = [sc.forward_device(f).clone() for f in sets]
        torch.cuda.synchronize()
        static = {k: v.clone() for k, v in sets[0].items()}
        out = torch.empty((B, 2), dtype=torch.float32, device="cuda")
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            sc.forward_device(static, probs_out=out)          # warm-up on the capture stream
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                sc.forward_device(static, probs_out=out)
        for rounds in range(2):
            for f, want in zip(sets, eager):
                for k in static:
                    static[k].copy_(f[k])
                g.replay()
                torch.cuda.synchronize()
                assert torch.equal(out, want)
    finally:
        sc.close()
