"""CPU tests of the rows SURVEY.md section 8f marks "next": record decode (N1, host C++ through the C ABI -- no GPU
involved), WordPiece tokenisation (N2, pinned by the reference's own tokenizer classes) and checkpoint name mapping (N3)."""
import base64
import json
import os

import numpy as np
import pytest
import torch

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import checkpoints, records, synth, tokenizer
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LXMERT, ZK, ModelConfig

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------------------------------------ N2 tokenizer
def _kat():
    with open(os.path.join(GOLD, "tokenizer_kat.json"), encoding="utf-8") as f:
        return json.load(f)


def test_tokenizer_matches_reference_classes_on_golden():
    """tests/golden/tokenizer_kat.json = outputs of imagebert_zk/tokenization.py (FullTokenizer) and
    lxmert/src/lxrt/tokenization.py (BertTokenizer) themselves (tools/make_golden.py --tokenizer)."""
    k = _kat()
    vocab = {t: i for i, t in enumerate(k["vocab"])}
    tf_tok = tokenizer.FullTokenizer(vocab=vocab, do_lower_case=True)                                  # 200-char limit
    lx_tok = tokenizer.FullTokenizer(vocab=vocab, do_lower_case=True, max_input_chars_per_word=100)    # lxmert's
    assert len(k["cases"]) >= 15
    for c in k["cases"]:
        a = tf_tok.tokenize(c["text"])
        assert a == c["tf_tokens"], c["text"]
        assert tf_tok.convert_tokens_to_ids(a) == c["tf_ids"]
        b = lx_tok.tokenize(c["text"])
        assert b == c["lxmert_tokens"], c["text"]
        assert lx_tok.convert_tokens_to_ids(b) == c["lxmert_ids"]
    assert tf_tok.convert_ids_to_tokens(k["cases"][0]["tf_ids"]) == k["cases"][0]["tf_tokens"]
    with pytest.raises(KeyError):
        tf_tok.convert_tokens_to_ids(["not-in-vocab"])


def test_mirror_module_exposes_reference_names():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code.imagebert_zk import tokenization
    assert tokenization.whitespace_tokenize("  a  b ") == ["a", "b"] and tokenization.whitespace_tokenize("") == []
    assert {"FullTokenizer", "BasicTokenizer", "WordpieceTokenizer", "load_vocab"} <= set(dir(tokenization))


# ------------------------------------------------------------------------------------------------ N1 decode
def _make_lines(n, rng, max_nb=14, bad=None):
    lines, want = [], []
    for i in range(n):
        nb = int(rng.integers(1, max_nb + 1))
        h, w = int(rng.integers(200, 900)), int(rng.integers(200, 900))
        boxes = (rng.random((nb, 4)) * 500).astype(np.float32)
        feats = rng.standard_normal((nb, 2048)).astype(np.float32)
        labels = rng.integers(0, 33, nb).astype(np.int64)
        query = ["women's leather shoes", "forest style 连衣裙", "kids wash basin"][i % 3] + f" {i}"
        fields = [str(1000 + i), str(h), str(w), str(nb), base64.b64encode(boxes.tobytes()).decode(),
                  base64.b64encode(feats.tobytes()).decode(), base64.b64encode(labels.tobytes()).decode(), query, str(7 * i)]
        lines.append(("\t".join(fields) + "\n").encode("utf-8"))
        want.append((1000 + i, h, w, nb, boxes, feats, labels, query, 7 * i))
    return lines, want


def _reference_read_line(line):
    """The decode part of read_line (imagebert_zk/load_data_v4.py:133-147), verbatim semantics."""
    arr = line.decode("utf-8").strip().split("\t")
    nb = int(arr[3])
    boxes = np.frombuffer(base64.b64decode(arr[4]), dtype=np.float32).reshape(nb, 4)
    feats = np.frombuffer(base64.b64decode(arr[5]), dtype=np.float32).reshape(nb, 2048)
    labels = np.frombuffer(base64.b64decode(arr[6]), dtype=np.int64).reshape(nb)
    return int(arr[0]), int(arr[1]), int(arr[2]), nb, boxes, feats, labels, arr[7], int(arr[8])


@pytest.mark.parametrize("threads", [1, 4])
def test_decode_matches_reference_read_line_bit_exact(threads):
    rng = np.random.default_rng(3)
    lines, _ = _make_lines(23, rng)
    R = 10
    out = records.decode_lines(lines, max_boxes=R, n_threads=threads, pin=False)
    for i, line in enumerate(lines):
        pid, h, w, nb, boxes, feats, labels, query, qid = _reference_read_line(line)
        k = min(nb, R)                                            # seq_padding_2: truncate to the box budget, zero pad
        assert (out["product_id"][i], out["image_h"][i], out["image_w"][i], out["num_boxes"][i], out["query_id"][i]) \
            == (pid, h, w, nb, qid)
        assert out["queries"][i] == query
        assert np.array_equal(out["boxes4"][i, :k].numpy(), boxes[:k]) and not out["boxes4"][i, k:].any()
        assert np.array_equal(out["feats"][i, :k].numpy().view(np.uint32), feats[:k].view(np.uint32))
        assert not out["feats"][i, k:].any()
        assert np.array_equal(out["class_labels"][i, :k].numpy(), labels[:k]) and not out["class_labels"][i, k:].any()


def test_decode_and_feed_assembly_match_the_reference_loader_golden(tmp_path):
    """tests/golden/records_kat.npz = outputs of the reference's OWN read_line / seq_padding / seq_padding_2 and label
    cleaning loop (code/imagebert_zk/load_data_v4.py, extracted by name by tools/make_golden.py --records, run with the
    reference's own tokenizer) on 8 synthetic TSV lines, one of them over the 10-box budget; both values of the
    sen2forest flag.  The C++ decoder, the tokenizer, the label-phrase table and the feed assembly against it."""
    g = np.load(os.path.join(GOLD, "records_kat.npz"))
    lines = [str(x).encode("utf-8") for x in g["lines"]]
    out = records.decode_lines(lines, max_boxes=10, n_threads=2, pin=False)
    (tmp_path / "multimodal_labels.txt").write_text("\n".join(str(x) for x in g["label_lines"]) + "\n", encoding="utf-8")
    label_map = records.load_label_map(str(tmp_path / "multimodal_labels.txt"))
    tok = tokenizer.FullTokenizer(vocab={str(t): i for i, t in enumerate(g["vocab"])})
    cfg = ModelConfig("imagebert_lds", n_layers=1, lq=20, nbox=10, vocab=len(g["vocab"]))   # (zk adds boxes: GPU test)
    feeds = {f: records.FeedAssembler(cfg, tok, label_map, sen2forest=bool(f)).assemble(out) for f in (0, 1)}
    for i in range(len(lines)):
        pid, h, w, nb, qid = g[f"s2f0_{i}_scalars"].tolist()
        k = min(nb, 10)
        assert (out["product_id"][i], out["image_h"][i], out["image_w"][i], out["num_boxes"][i], out["query_id"][i]) \
            == (pid, h, w, nb, qid)
        assert np.array_equal(out["feats"][i, :k].numpy().view(np.uint32), g[f"s2f0_{i}_feats_u32"][:k])
        # raw boxes: the reference only keeps the normalised ones (float32 boxes / python ints = a float64 division)
        b4 = out["boxes4"][i, :k].numpy().astype(np.float64)
        want5 = g[f"s2f0_{i}_boxes5"][:k]
        assert np.array_equal((b4 / [h, w, h, w]).astype(np.float32), want5[:, :4])
        for f in (0, 1):
            ids = g[f"s2f{f}_{i}_query_ids"].tolist()
            assert feeds[f]["query_ids"][i].tolist() == (ids + [0] * 20)[:20]
            assert feeds[f]["label_ids"][i, :k].tolist() == g[f"s2f{f}_{i}_label_ids"][:k].tolist()
            assert not feeds[f]["label_ids"][i, k:].any()
    assert str(g["s2f1_1_query"]) == "forest style dress" and str(g["s2f0_1_query"]) == "sen department of dress"
    assert feeds[1]["query_ids"][1].tolist() != feeds[0]["query_ids"][1].tolist()
    # the zero-padded feature block of the feeds (seq_padding_2 to the box budget, load_data_v4.py:380-383)
    assert np.array_equal(out["feats"].numpy().view(np.uint32), g["feats_padded_u32"])


def test_lxmert_feeds_match_the_reference_loader_golden(tmp_path):
    """tests/golden/records_kat_lxmert.npz = the same TSV lines through the LXMERT tree's OWN read_line / seq_padding /
    seq_padding_2 (code/lxmert/src/utils.py, extracted by name) with its own BertTokenizer: query ids and mask at 23
    tokens, label-phrase ids, visual mask, and the 4-d boxes as float32(float64 division)."""
    g0 = np.load(os.path.join(GOLD, "records_kat.npz"))
    g = np.load(os.path.join(GOLD, "records_kat_lxmert.npz"))
    lines = [str(x).encode("utf-8") for x in g0["lines"]]
    out = records.decode_lines(lines, max_boxes=10, n_threads=1, pin=False)
    (tmp_path / "labels.txt").write_text("\n".join(str(x) for x in g0["label_lines"]) + "\n", encoding="utf-8")
    tok = tokenizer.FullTokenizer(vocab={str(t): i for i, t in enumerate(g0["vocab"])}, max_input_chars_per_word=100)
    cfg = ModelConfig(LXMERT, n_layers=1, n_r_layers=1, n_x_layers=1, lq=23, nbox=10, vocab=len(g0["vocab"]))
    fa = records.FeedAssembler(cfg, tok, records.load_label_map(str(tmp_path / "labels.txt")))
    # everything of assemble() but the GPU box normalisation (tests/test_gpu_records.py covers that one)
    orig = records.normalize_boxes
    records.normalize_boxes = lambda b4, h, w, with_area, device=None: torch.from_numpy(
        (b4.numpy().astype(np.float64) / np.stack([h.numpy(), w.numpy(), h.numpy(), w.numpy()], 1)[:, None, :]).astype(np.float32))
    try:
        feeds = fa.assemble(out)
    finally:
        records.normalize_boxes = orig
    assert np.array_equal(feeds["query_ids"].numpy(), g["query_ids_padded"])
    assert np.array_equal(feeds["query_mask"].numpy(), g["query_mask"])
    assert np.array_equal(feeds["visn_mask"].numpy(), g["visn_mask"])
    assert np.array_equal(feeds["boxes"].numpy(), g["boxes4_padded_f32"])
    for i in range(len(lines)):
        k = min(int(out["num_boxes"][i]), 10)
        assert feeds["label_ids"][i, :k].tolist() == g[f"{i}_label_ids"][:k].tolist()
        assert (out["product_id"][i], out["query_id"][i]) == tuple(g[f"{i}_ids"].tolist())


def test_decode_edge_cases_and_errors():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200._lib import MmrError
    assert records.decode_lines([], pin=False)["queries"] == []
    rng = np.random.default_rng(4)
    lines, _ = _make_lines(3, rng)
    cols = lines[1].split(b"\t")
    with pytest.raises(MmrError, match="line 1: fewer than 9"):
        records.decode_lines([lines[0], b"\t".join(cols[:6]), lines[2]], pin=False)
    broken = list(cols)
    broken[5] = broken[5][:-8]                                   # feature blob too short for num_boxes
    with pytest.raises(MmrError, match="base64"):
        records.decode_lines([b"\t".join(broken)], pin=False)
    broken = list(cols)
    broken[3] = b"x3"
    with pytest.raises(MmrError, match="integer"):
        records.decode_lines([b"\t".join(broken)], pin=False)
    # CRLF line endings and a missing final newline are accepted (.strip())
    ok = records.decode_lines([lines[0].rstrip(b"\n") + b"\r\n", lines[2].rstrip(b"\n")], pin=False)
    assert len(ok["queries"]) == 2


def test_reused_decoder_arrays_equal_a_fresh_decode():
    """RecordDecoder clears only the box slots an earlier batch left non-zero (mmr_decode_tsv_reuse): whatever was
    decoded into the arrays before -- more boxes, fewer lines, a batch that failed half-way, concurrent callers of the
    thread pool -- the arrays must be byte-identical to a one-shot decode of the same lines."""
    import threading

    from kddcup_2020_multimodalitiesrecall_2nd_place_b200._lib import MmrError
    rng = np.random.default_rng(11)
    R = 12
    big, _ = _make_lines(9, rng, max_nb=20)               # some records over the box budget
    small, _ = _make_lines(9, rng, max_nb=3)
    mid, _ = _make_lines(5, rng, max_nb=9)
    dec = records.RecordDecoder(9, max_boxes=R, n_threads=3, pin=False)
    keys = ("product_id", "image_h", "image_w", "num_boxes", "boxes4", "feats", "class_labels", "query_id")

    def check(lines):
        got = dec.decode(lines)
        want = records.decode_lines(lines, max_boxes=R, n_threads=1, pin=False)
        for k in keys:
            assert torch.equal(got[k], want[k]), k
        assert got["queries"] == want["queries"]

    for lines in (big, small, mid, big[::-1], small):
        check(lines)
    cols = big[2].split(b"\t")
    cols[5] = cols[5][:-8]                                  # fails after its boxes and part of its features were written
    with pytest.raises(MmrError, match="base64"):
        dec.decode([big[0], big[1], b"\t".join(cols)])
    check(small)
    check(mid)
    # several decoders at once on the shared worker pool
    errs = []

    def worker(seed):
        try:
            r = np.random.default_rng(seed)
            d = records.RecordDecoder(6, max_boxes=R, n_threads=4, pin=False)
            for _ in range(4):
                ls, _ = _make_lines(int(r.integers(1, 7)), r, max_nb=14)
                got = d.decode(ls)
                want = records.decode_lines(ls, max_boxes=R, n_threads=1, pin=False)
                assert all(torch.equal(got[k], want[k]) for k in keys)
        except Exception as e:                              # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_feed_assembly_follows_the_loaders():
    """Token ids / lengths / masks / label phrases as load_data_v4.py:148-163, 204, 259-265 and utils.py:38-59 build them."""
    k = _kat()
    vocab = {t: i for i, t in enumerate(k["vocab"])}
    tok = tokenizer.FullTokenizer(vocab=vocab)
    label_map = {0: "women dress", 1: "leather shoes", 2: "kids", 5: "wash basin"}
    cfg = ModelConfig(LXMERT, n_layers=1, n_r_layers=1, n_x_layers=1, lq=12, nbox=4, vocab=len(k["vocab"]))
    fa = records.FeedAssembler(cfg, tok, label_map)
    batch = {"queries": ["women leather shoes", "kids"], "num_boxes": torch.tensor([2, 6], dtype=torch.int32),
             "class_labels": torch.tensor([[1, 5, 0, 0], [2, 0, 1, 5]]), "feats": torch.zeros(2, 4, 2048)}
    cfg_lds = ModelConfig("imagebert_lds", n_layers=1, lq=12, nbox=4, vocab=len(k["vocab"]))
    feeds = records.FeedAssembler(cfg_lds, tok, label_map).assemble(batch)
    ids = lambda s: tok.convert_tokens_to_ids(tok.tokenize(s))
    want_q0 = [vocab["[CLS]"]] + ids("women leather shoes") + [vocab["[SEP]"]]
    assert feeds["query_ids"][0, :len(want_q0)].tolist() == want_q0 and not feeds["query_ids"][0, len(want_q0):].any()
    assert feeds["label_ids"][0, 0, :2].tolist() == ids("leather shoes") and not feeds["label_ids"][0, 2:].any()
    assert feeds["label_ids"][1, 3, :2].tolist() == ids("wash basin")          # 6 boxes in the file, 4 slots kept
    assert feeds["segment_ids"].shape == (2, 12) and not feeds["segment_ids"].any()
    assert fa.label_table.shape == (6, 8)


# ------------------------------------------------------------------------------------------------ N3 checkpoints
def test_tf_checkpoint_selection_prefers_ema_and_drops_optimizer_slots(tmp_path):
    cfg = ModelConfig(ZK, n_layers=1, lq=20, nbox=10, vocab=50)
    w = synth.make_weights(cfg, seed=1)
    entries = {}
    for k_, v in w.items():
        entries[k_] = v + 1.0                                    # raw variable (must lose against its shadow)
        entries[k_ + "/ExponentialMovingAverage"] = v
        entries[k_ + "/adam_m"] = np.zeros_like(v)
        entries[k_ + "/adam_v"] = np.zeros_like(v)
    entries["global_step"] = np.array(1234, np.int64)
    got, src = checkpoints.select_tf_variables(entries, prefer_ema=True, wanted=w.keys())
    assert set(got) == set(w) and all(np.array_equal(got[k_], w[k_]) for k_ in w)
    assert all(s.endswith("/ExponentialMovingAverage") for s in src.values())
    raw, _ = checkpoints.select_tf_variables(entries, prefer_ema=False)
    assert np.array_equal(raw["bert/pooler/dense/bias"], w["bert/pooler/dense/bias"] + 1.0) and "global_step" not in raw
    np.savez(tmp_path / "ckpt.npz", **{k_.replace("/", "|"): v for k_, v in list(entries.items())[:3]})
    with pytest.raises(KeyError, match="checkpoint lacks"):
        checkpoints.select_tf_variables({"a": np.zeros(1)}, wanted=["b"])


def test_pth_import_strips_dataparallel_prefix(tmp_path):
    cfg = ModelConfig(LXMERT, n_layers=1, n_r_layers=1, n_x_layers=1, vocab=40)
    w = synth.make_weights(cfg, seed=2)
    sd = {"module." + k_: torch.from_numpy(v).double() for k_, v in w.items()}
    sd["module.some.counter"] = torch.tensor(3)
    torch.save(sd, tmp_path / "BEST.pth")
    got = checkpoints.load_pth(str(tmp_path / "BEST"))           # the reference passes the path without ".pth"
    assert set(got) == set(w) and all(got[k_].dtype == np.float32 and np.array_equal(got[k_], w[k_]) for k_ in w)
