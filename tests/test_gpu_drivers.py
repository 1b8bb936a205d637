"""End-to-end drivers on the GPU at small size: TSV file -> C++ decode -> feeds -> three scorers -> the four score files
in the reference's formats -> code/main.py ensemble -> submission.csv; KDD.load / KDD.predict (kdd_model.py:46-152) from
a `.pth`; zk from a TF checkpoint on disk read without TensorFlow.  Scores are compared with the fp32 oracle fed the
reference's way (the per-line arithmetic of read_line, restated in test_gpu_records.py)."""
import csv
import json
import os

import numpy as np
import pytest
import torch

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import (checkpoints, drivers, ensemble, records, synth, tf_bundle,
                                                              tokenizer)
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LDS, LXMERT, ZK, ModelConfig
from tests.test_gpu_records import _lines

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LABELS = {0: "women dress", 1: "leather shoes", 2: "kids", 3: "wash basin", 4: "black shirt", 5: "men"}
QUERIES = ["women's leather shoes", "sen department of dress 女士", "kids wash basin red", "running shoes for men"]


def _vocab():
    k = json.load(open(os.path.join(GOLD, "tokenizer_kat.json"), encoding="utf-8"))
    return {t: i for i, t in enumerate(dict.fromkeys(k["vocab"]))}      # (the KAT list repeats two tokens)


def _write_tsv(path, n, seed=5):
    lines = _lines(n, 13, np.random.default_rng(seed), QUERIES)
    with open(path, "wb") as f:
        f.write(b"product_id\timage_h\timage_w\tnum_boxes\tboxes\tfeatures\tclass_labels\tquery\tquery_id\n")
        f.writelines(lines)
    return lines


def test_tsv_to_submission_through_three_models(tmp_path):
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import MatchScorer
    vocab = _vocab()
    n = 45                                                      # 15 queries x 3 candidates; chunks of 16: ragged tail
    tsv = str(tmp_path / "testB.tsv")
    lines = _write_tsv(tsv, n)
    assert drivers.read_tsv_lines(tsv) == lines                  # header skipped
    shapes = dict(lq=20, nbox=10, vocab=max(vocab.values()) + 1)
    cfgs = {ZK: ModelConfig(ZK, n_layers=2, **shapes), LDS: ModelConfig(LDS, n_layers=2, **shapes),
            LXMERT: ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=1, **shapes)}
    scorers = {k: MatchScorer(c, synth.make_weights(c, seed=31, trained_like=True), device=0, max_batch=16)
               for k, c in cfgs.items()}
    toks = {k: tokenizer.FullTokenizer(vocab=vocab, max_input_chars_per_word=100 if k == LXMERT else 200) for k in cfgs}
    try:
        rows = drivers.run_ensemble(scorers, toks, LABELS, tsv, str(tmp_path / "prediction_result"))
        # per-model files exist in the reference's formats and reproduce the submission through code/main.py itself
        out = tmp_path / "prediction_result"
        zk = ensemble.read_scores(str(out / "testB_result_match_keyword_valid_finetune_251.txt"))
        s2f = ensemble.read_scores(str(out / "testB_result_match_keyword_valid_finetune_251_sen_to_forest.txt"))
        lx = ensemble.read_scores(str(out / "testB_score_lxmert.csv"), ",", "query")
        assert sum(len(v) for v in zk.values()) == n and sum(len(v) for v in lx.values()) == n
        assert open(out / "testB_score_lxmert.csv").readline().strip() == "query-id,product-id,score"
        # the sen2forest rewrite changes the scores of exactly the queries that contain the phrase (queries 1, 5, 9, ...)
        changed = {q for q in zk for p in zk[q] if zk[q][p] != s2f[q][p]}
        ids = {str(i // 3) for i in range(n) if "sen department of" in QUERIES[i % 4]}
        assert changed and changed <= ids
        with open(out / "submission.csv") as f:
            got = list(csv.reader(f))
        assert got[0] == ["query-id", "product1", "product2", "product3", "product4", "product5"]
        assert [r[0] for r in got[1:]] == [q for q, _ in rows]
        # scores against one direct scorer call on the same lines (no chunking, no double buffering)
        batch = records.decode_lines(lines, max_boxes=10)
        feeds = records.FeedAssembler(cfgs[ZK], toks[ZK], LABELS).assemble(batch)
        big = MatchScorer(cfgs[ZK], synth.make_weights(cfgs[ZK], seed=31, trained_like=True), device=0, max_batch=n)
        want = big.score({k: v.cpu() for k, v in feeds.items()})[:, 1].numpy()
        big.close()
        have = np.array([zk[str(i // 3)][str(i)] for i in range(n)], np.float32)
        assert np.array_equal(have, want)
    finally:
        for sc in scorers.values():
            sc.close()
    with pytest.raises(KeyError, match="class id"):
        bad = dict(LABELS)
        del bad[3]
        sc = MatchScorer(cfgs[LDS], synth.make_weights(cfgs[LDS], seed=31), device=0, max_batch=16)
        try:
            drivers.score_tsv(sc, toks[LDS], bad, lines)
        finally:
            sc.close()


def test_kdd_load_and_predict_from_pth(tmp_path):
    """lxmert: KDD.load(path) reads `<path>.pth` (DataParallel prefixes), KDD.predict(mod, save=True) scores
    data/<mod>/<mod>.tsv and writes <result>/<mod>_score_lxmert.csv (kdd_model.py:46-152)."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code.lxmert.src.tasks.kdd_model import KDD
    from oracle import imagebert, lxmert
    vocab = _vocab()
    (tmp_path / "vocab.txt").write_text("\n".join(sorted(vocab, key=vocab.get)) + "\n", encoding="utf-8")
    (tmp_path / "labels.txt").write_text("".join(f"{i}\t{p}\n" for i, p in LABELS.items()), encoding="utf-8")
    (tmp_path / "data" / "valid").mkdir(parents=True)
    lines = _write_tsv(str(tmp_path / "data" / "valid" / "valid.tsv"), 20, seed=9)
    cfg = ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=1, lq=23, nbox=10, vocab=max(vocab.values()) + 1)
    w = synth.make_weights(cfg, seed=41)
    torch.save({"module." + k: torch.from_numpy(v) for k, v in w.items()}, str(tmp_path / "BEST.pth"))
    kdd = KDD(str(tmp_path / "data"), str(tmp_path / "result"), str(tmp_path / "vocab.txt"), str(tmp_path / "labels.txt"),
              batch_size=8)
    kdd.load(str(tmp_path / "BEST"))
    try:
        match_pred, match_label, rank = kdd.predict("valid", save=True)
    finally:
        kdd.scorer.close()
    assert len(match_pred) == 20 and match_label == [0] * 20 and sum(len(v) for v in rank.values()) == 20
    saved = ensemble.read_scores(str(tmp_path / "result" / "valid_score_lxmert.csv"), ",", "query")
    assert all(abs(saved[str(q)][str(p)] - s) < 1e-7 for q, v in rank.items() for p, s in v)
    # oracle on the same records, fed the reference's way
    tok = tokenizer.FullTokenizer(vocab=vocab, max_input_chars_per_word=100)
    batch = records.decode_lines(lines, max_boxes=10)
    feeds = records.FeedAssembler(cfg, tok, LABELS).assemble(batch)
    inp = {k: v.cpu() for k, v in feeds.items()}
    ref = lxmert.forward(imagebert.to_torch(w), inp, 2, 1, 1)["probs"][:, 1].numpy()
    got = np.array([dict(rank[i // 3])[i] for i in range(20)], np.float32)
    assert np.abs(got - ref).max() <= 1e-3


def test_zk_scores_from_a_tf_checkpoint_on_disk(tmp_path):
    """evaluate_normal.py:204-212 without TensorFlow: a V2 checkpoint (EMA shadows + raw variables + global_step) ->
    tf_bundle -> EMA selection -> scorer; same scores as the scorer built from the dict directly."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import MatchScorer
    cfg = ModelConfig(ZK, n_layers=2, lq=20, nbox=8, vocab=500)
    w = synth.make_weights(cfg, seed=51)
    stored = {"global_step": np.array(251, np.int64)}
    for name, a in w.items():
        stored[name] = a * 0.5                                   # the raw variable: must NOT be used
        stored[name + "/ExponentialMovingAverage"] = a
    (tmp_path / "ck").mkdir()
    tf_bundle.write_bundle(str(tmp_path / "ck" / "model.ckpt-251"), stored)
    (tmp_path / "ck" / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-251"\n')
    loaded = checkpoints.tf_variables_from_checkpoint(str(tmp_path / "ck"), prefer_ema=True, verify_data=True)
    inp = synth.make_inputs(cfg, 6, seed=51)
    out = []
    for weights in (w, loaded):
        sc = MatchScorer(cfg, weights, device=0, max_batch=6)
        out.append(sc.score(sc.to_feeds(inp)).clone())
        sc.close()
    assert torch.equal(out[0], out[1])
