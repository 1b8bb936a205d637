"""TSV record -> decode (C++) -> feeds (tokenizer, label phrases, GPU box normalisation) -> scorer, against the oracle fed
with what the reference's own read_line / get_batch arithmetic produces for the same lines."""
import base64
import json
import os

import numpy as np
import pytest
import torch

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import records, synth, tokenizer
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LXMERT, ZK, ModelConfig

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _lines(n, R_file, rng, queries):
    out = []
    for i in range(n):
        nb = int(rng.integers(1, R_file + 1))
        h, w = int(rng.integers(200, 900)), int(rng.integers(200, 900))
        x = np.sort(rng.random((nb, 2, 2)).astype(np.float32), axis=1) * np.array([h, w], np.float32)
        boxes = np.stack([x[:, 0, 0], x[:, 0, 1], x[:, 1, 0], x[:, 1, 1]], 1).astype(np.float32)
        feats = (np.abs(rng.standard_normal((nb, 2048))) * 0.5 * (rng.random((nb, 2048)) > 0.6)).astype(np.float32)
        labels = rng.integers(0, 6, nb).astype(np.int64)
        f = [str(i), str(h), str(w), str(nb), base64.b64encode(boxes.tobytes()).decode(),
             base64.b64encode(feats.tobytes()).decode(), base64.b64encode(labels.tobytes()).decode(),
             queries[i % len(queries)], str(i // 3)]
        out.append(("\t".join(f) + "\n").encode())
    return out


def test_boxes_normalize_matches_numpy_float64_division():
    rng = np.random.default_rng(0)
    n, R = 37, 10
    b4 = (rng.random((n, R, 4)) * 800).astype(np.float32)
    h = rng.integers(100, 1000, n).astype(np.int32)
    w = rng.integers(100, 1000, n).astype(np.int32)
    # load_data_v4.py:142-145: float32 boxes / python list of ints -> float64, stored into a float32 array
    want5 = np.zeros((n, R, 5), np.float32)
    for i in range(n):
        want5[i, :, :4] = b4[i] / [h[i], w[i], h[i], w[i]]
        want5[i, :, 4] = (b4[i, :, 2] - b4[i, :, 0]) * (b4[i, :, 3] - b4[i, :, 1]) / (w[i] * h[i])
    got5 = records.normalize_boxes(torch.from_numpy(b4), torch.from_numpy(h), torch.from_numpy(w), with_area=True).cpu().numpy()
    got4 = records.normalize_boxes(torch.from_numpy(b4), torch.from_numpy(h), torch.from_numpy(w), with_area=False).cpu().numpy()
    assert np.array_equal(got5, want5) and np.array_equal(got4, want5[..., :4])


def test_zk_feeds_match_the_reference_loader_golden(tmp_path):
    """The zk feeds of FeedAssembler.assemble — GPU box normalisation + area included — against what the reference's
    own read_line / seq_padding_2 produce for the same TSV lines (tests/golden/records_kat.npz): bit-exact."""
    g = np.load(os.path.join(GOLD, "records_kat.npz"))
    lines = [str(x).encode("utf-8") for x in g["lines"]]
    out = records.decode_lines(lines, max_boxes=10, pin=False)
    (tmp_path / "labels.txt").write_text("\n".join(str(x) for x in g["label_lines"]) + "\n", encoding="utf-8")
    tok = tokenizer.FullTokenizer(vocab={str(t): i for i, t in enumerate(g["vocab"])})
    cfg = ModelConfig(ZK, n_layers=1, lq=20, nbox=10, vocab=len(g["vocab"]))
    feeds = records.FeedAssembler(cfg, tok, records.load_label_map(str(tmp_path / "labels.txt"))).assemble(out)
    assert np.array_equal(feeds["boxes"].cpu().numpy(), g["boxes5_padded"])
    for i in range(len(lines)):
        ids = g[f"s2f0_{i}_query_ids"].tolist()
        assert feeds["len_query"][i].item() == min(len(ids), 20)
        assert feeds["num_boxes"][i].item() == min(int(g[f"s2f0_{i}_scalars"][3]), 10)


@pytest.mark.parametrize("kind", [ZK, LXMERT])
def test_tsv_to_scores_against_oracle(kind):
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import MatchScorer
    from oracle import imagebert, lxmert
    k = json.load(open(os.path.join(GOLD, "tokenizer_kat.json"), encoding="utf-8"))
    vocab = {t: i for i, t in enumerate(k["vocab"])}
    tok = tokenizer.FullTokenizer(vocab=vocab, max_input_chars_per_word=100 if kind == LXMERT else 200)
    label_map = {0: "women dress", 1: "leather shoes", 2: "kids", 3: "wash basin", 4: "black shirt", 5: "men"}
    queries = ["women's leather shoes", "forest style dress 女士", "kids wash basin red", "running shoes for men"]
    R, Lq, n = 10, 20, 12
    if kind == ZK:
        cfg = ModelConfig(ZK, n_layers=2, lq=Lq, nbox=R, vocab=len(k["vocab"]))
    else:
        cfg = ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=1, lq=Lq, nbox=R, vocab=len(k["vocab"]))
    lines = _lines(n, 13, np.random.default_rng(5), queries)           # some records exceed the 10-box budget
    batch = records.decode_lines(lines, max_boxes=R)
    feeds = records.FeedAssembler(cfg, tok, label_map).assemble(batch)
    w = synth.make_weights(cfg, seed=21)
    sc = MatchScorer(cfg, w, device=0, max_batch=n)
    got = sc.score({k_: (v.cpu() if torch.is_tensor(v) else v) for k_, v in feeds.items()})[:, 1]
    # oracle inputs built the reference's way, line by line (read_line + seq_padding + get_batch bookkeeping)
    inp = {kk: [] for kk in ("query_ids", "len_query", "num_boxes", "feats", "label_ids", "boxes")}
    for line in lines:
        arr = line.decode().strip().split("\t")
        nb, h, wd = int(arr[3]), int(arr[1]), int(arr[2])
        boxes = np.frombuffer(base64.b64decode(arr[4]), dtype=np.float32).reshape(nb, 4)
        b5 = np.zeros((nb, 5), np.float32)
        b5[:, :4] = boxes / [h, wd, h, wd]
        b5[:, 4] = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) / (wd * h)
        feats = np.frombuffer(base64.b64decode(arr[5]), dtype=np.float32).reshape(nb, 2048)
        cls = np.frombuffer(base64.b64decode(arr[6]), dtype=np.int64)
        lab = [(tok.convert_tokens_to_ids(tok.tokenize(label_map[int(c)])) + [0] * 8)[:8] for c in cls]
        q = tok.convert_tokens_to_ids(["[CLS]"] + tok.tokenize(arr[7]) + ["[SEP]"])
        pad2 = lambda x, c: np.concatenate([x[:R], np.zeros((max(0, R - len(x)), c), x.dtype)])
        inp["query_ids"].append((q + [0] * Lq)[:Lq])
        inp["len_query"].append(min(len(q), Lq))
        inp["num_boxes"].append(min(nb, R))
        inp["feats"].append(pad2(feats, 2048))
        inp["label_ids"].append(pad2(np.array(lab, np.int32), 8))
        inp["boxes"].append(pad2(b5 if kind == ZK else b5[:, :4].copy(), 5 if kind == ZK else 4))
    inp = {kk: np.array(v) for kk, v in inp.items()}
    if kind == ZK:
        inp["segment_ids"] = np.tile(np.array([0] * Lq + [1] * R, np.int32), (n, 1))
        inp["labels"] = np.ones(n, np.int32)
        ref = imagebert.zk_forward(imagebert.to_torch(w), imagebert.to_torch(inp), cfg.n_layers)["probs"][:, 1]
    else:
        inp["query_mask"] = (np.arange(Lq)[None] < inp["len_query"][:, None]).astype(np.int32)
        inp["visn_mask"] = (np.arange(R)[None] < inp["num_boxes"][:, None]).astype(np.int32)
        ref = lxmert.forward(imagebert.to_torch(w), imagebert.to_torch(inp), 2, 1, 1)["probs"][:, 1]
    assert (got - ref).abs().max().item() <= 1e-3
    sc.close()
