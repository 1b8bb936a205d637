"""Generates the committed golden fixtures under tests/golden/ from the reference tree (dev container only).

    python tools/make_golden.py

  ensemble_kat.npz  the four shipped per-model score files + the shipped submission.csv (code/main.py KAT)
  ndcg_kat.npz      shipped valid scores + valid answers (nDCG@5 known answer 0.7098, report table 5)
  lxmert_ref_*.npz  outputs of the reference's OWN LXMERT code (oracle/lxmert_ref.py) on seeded synthetic
                    weights/inputs from kddcup_2020_multimodalitiesrecall_2nd_place_b200.synth

    python tools/make_golden.py --tf-shim [--check]

  zk_ref_shim_*.npz, lds_ref_shim_*.npz   outputs of the reference's OWN TF-1 model code (imagebert_zk/pixelbert.py +
                    model_triple.py; imagebert_lds/src/pixelmodel.py + get_next_sentence_output of
                    run_pretraining_predict_score.py) executed unmodified on tools/tf1_shim.py, an eager stand-in for
                    the TensorFlow ops it calls (TF 1.12 itself cannot be installed here).  --check regenerates in
                    memory and compares with the committed files instead of writing them.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def read_scores(path, sep, header=False):
    q, p, s = [], [], []
    for line in open(path):
        if header and "query" in line:
            continue
        a = line.strip().split(sep)
        q.append(int(a[0])); p.append(int(a[1])); s.append(float(a[2]))
    return np.array(q, np.int64), np.array(p, np.int64), np.array(s, np.float64)


def ensemble_kat():
    pr = os.path.join(REF, "prediction_result")
    files = [("zk", "testB_result_match_keyword_valid_finetune_251.txt", "\t", False),
             ("zk_s2f", "testB_result_match_keyword_valid_finetune_251_sen_to_forest.txt", "\t", False),
             ("lds", "testBscore_imagebert.txt", "\t", False),
             ("lxmert", "testB_score_lxmert.csv", ",", True)]
    out = {}
    for tag, fn, sep, hdr in files:
        q, p, s = read_scores(os.path.join(pr, fn), sep, hdr)
        out[tag + "_q"], out[tag + "_p"], out[tag + "_s"] = q, p, s
    rows = []
    for i, line in enumerate(open(os.path.join(pr, "submission.csv"))):
        if i == 0:
            continue
        rows.append([int(x) for x in line.strip().split(",")])
    out["submission"] = np.array(rows, np.int64)
    np.savez_compressed(os.path.join(OUT, "ensemble_kat.npz"), **out)
    print("ensemble_kat:", {k: v.shape for k, v in out.items()})


def ndcg_kat():
    q, p, s = read_scores(os.path.join(REF, "code/imagebert_lds/src/validscore_imagebert.txt"), "\t")
    ans_q, ans = [], []
    for i, line in enumerate(open(os.path.join(REF, "code/imagebert_zk/valid_answer.txt"))):
        if i == 0:
            continue
        a = [int(x) for x in line.strip().split("\t")]
        ans_q.append(a[0])
        ans.append(a[1:] + [-1] * (7 - len(a)))
    np.savez_compressed(os.path.join(OUT, "ndcg_kat.npz"), q=q, p=p, s=s, ans_q=np.array(ans_q, np.int64),
                        ans=np.array(ans, np.int64), expected=np.array(0.7098))
    print("ndcg_kat:", len(q), len(ans_q))


def weights_digest(w):
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode()); h.update(np.ascontiguousarray(w[k]).tobytes())
    return h.hexdigest()


def lxmert_ref():
    import torch
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import synth
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LXMERT, ModelConfig
    from oracle import lxmert_ref as ref
    torch.set_num_threads(8)
    cases = {
        "small": (ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=2, lq=23, nbox=10, vocab=2000), 4, False),
        "small_trained": (ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=2, lq=23, nbox=10, vocab=2000), 4, True),
        "native": (ModelConfig(LXMERT, n_layers=9, n_r_layers=5, n_x_layers=5, lq=23, nbox=10, vocab=2000), 3, True),
        "cfg3shape": (ModelConfig(LXMERT, n_layers=2, n_r_layers=2, n_x_layers=2, lq=32, nbox=36, vocab=2000), 2, True),
    }
    for tag, (cfg, B, tl) in cases.items():
        w = synth.make_weights(cfg, seed=synth.SEED0 + 3, trained_like=tl)
        inp = synth.make_inputs(cfg, B, seed=synth.SEED0 + 3)
        model = ref.build_reference_model(cfg, w)
        out = ref.reference_forward(model, inp)
        np.savez_compressed(
            os.path.join(OUT, f"lxmert_ref_{tag}.npz"), cfg=np.array(str(cfg.to_dict())), batch=np.array(B),
            trained_like=np.array(tl), seed=np.array(synth.SEED0 + 3), weights_sha256=np.array(weights_digest(w)),
            probs=out["probs"].numpy(), logit=out["logit"].numpy(), x_norm=out["x_norm"].numpy())
        print("lxmert_ref", tag, out["probs"][:, 1].numpy())


# ---------------------------------------------------------------------------------------------- TF-1 model code on the shim
def _bert_config_json(cfg, path):
    import json
    json.dump({"vocab_size": cfg.vocab, "hidden_size": cfg.hidden, "num_hidden_layers": cfg.n_layers,
               "num_attention_heads": cfg.heads, "intermediate_size": cfg.intermediate, "hidden_act": "gelu",
               "hidden_dropout_prob": 0.1, "attention_probs_dropout_prob": 0.1, "max_position_embeddings": cfg.max_pos,
               "type_vocab_size": cfg.type_vocab, "initializer_range": 0.02}, open(path, "w"))


def _fresh_import(name, src_dir):
    import importlib
    for m in ("model_triple", "pixelbert", "pixelmodel"):
        sys.modules.pop(m, None)
    if src_dir in sys.path:
        sys.path.remove(src_dir)
    sys.path.insert(0, src_dir)
    return importlib.import_module(name)


def zk_on_shim(cfg, w, inp):
    """model_triple.model_attention_channel_e (code/imagebert_zk/model_triple.py:162-214) as the reference's
    evaluate_normal.py:222-236 feeds it, is_training=False.  The reference hard-codes 20 query tokens and 10 boxes."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import tf1_shim
    tf1_shim.install(w)
    T = tf1_shim.Tensor
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "user_data"))
        os.makedirs(os.path.join(d, "work"))
        _bert_config_json(cfg, os.path.join(d, "user_data", "bert_config.json"))
        os.chdir(os.path.join(d, "work"))          # model_triple.py:19 reads '../user_data/bert_config.json' at import
        try:
            mt = _fresh_import("model_triple", os.path.join(REF, "code/imagebert_zk"))
            captured = {}
            orig = mt.image_bert

            def image_bert(*a, **k):
                captured["model"] = orig(*a, **k)
                return captured["model"]
            mt.image_bert = image_bert
            _, probs, _ = mt.model_attention_channel_e(
                T(inp["num_boxes"]), T(inp["boxes"]), T(inp["feats"]), T(inp["label_ids"]), None, T(inp["query_ids"]),
                T(inp["len_query"]), T(inp["labels"]), T(inp["segment_ids"]), None, None, is_training=False)
        finally:
            os.chdir(cwd)
    m = captured["model"]
    return {"probs": probs.numpy(), "pooled": m.get_pooled_output().numpy(),
            "embedding_output": m.embedding_output.numpy(), "sequence_output": m.sequence_output.numpy(),
            "variables": sorted(tf1_shim.used_variables())}


def lds_on_shim(cfg, w, inp):
    """pixelmodel.BertModel called as bertmodel() does (run_pretraining_predict_score.py:324-331) and that file's
    get_next_sentence_output (479-501), extracted by name (importing the script would run its flag definitions)."""
    import ast
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import tf1_shim
    tf = tf1_shim.install(w)
    T = tf1_shim.Tensor
    src = os.path.join(REF, "code/imagebert_lds/src")
    pm = _fresh_import("pixelmodel", src)
    with tempfile.TemporaryDirectory() as d:
        _bert_config_json(cfg, os.path.join(d, "bert_config.json"))
        bert_config = pm.BertConfig.from_json_file(os.path.join(d, "bert_config.json"))
    model = pm.BertModel(imgfeat=T(inp["feats"]), config=bert_config, is_training=False, input_ids=T(inp["query_ids"]),
                         label_ids=T(inp["label_ids"]), token_type_ids=T(inp["segment_ids"]),
                         use_one_hot_embeddings=False, random_sample=False)
    tree = ast.parse(open(os.path.join(src, "run_pretraining_predict_score.py")).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_next_sentence_output"][0]
    ns = {"tf": tf, "pixelmodel": pm}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "run_pretraining_predict_score.py", "exec"), ns)
    _, _, _, probs = ns["get_next_sentence_output"](bert_config, model.get_pooled_output(), T(inp["labels"]))
    return {"probs": probs.numpy(), "pooled": model.get_pooled_output().numpy(),
            "embedding_output": model.embedding_output.numpy(), "sequence_output": model.sequence_output.numpy(),
            "variables": sorted(tf1_shim.used_variables())}


def _token_stats(x):
    """[B, S, H] activations -> per-token mean and L2 norm (small enough to commit)."""
    return np.stack([x.mean(-1), np.sqrt((x * x).sum(-1))], -1).astype(np.float32)


def tf_shim_refs(check=False):
    import torch
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import synth
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LDS, ZK, ModelConfig
    torch.set_num_threads(8)
    cases = {"small": (2, 4, False), "small_trained": (2, 4, True), "native": (12, 3, True)}
    worst = 0.0
    for kind, fn in ((ZK, zk_on_shim), (LDS, lds_on_shim)):
        for tag, (layers, B, tl) in cases.items():
            cfg = ModelConfig(kind, n_layers=layers, lq=20, nbox=10, vocab=2000)     # the reference's native shapes
            seed = synth.SEED0 + 5
            w = synth.make_weights(cfg, seed=seed, trained_like=tl)
            inp = synth.make_inputs(cfg, B, seed=seed)
            out = fn(cfg, w, inp)
            assert out["variables"] == sorted(w), (sorted(set(w) ^ set(out["variables"])))
            rec = dict(cfg=np.array(str(cfg.to_dict())), batch=np.array(B), trained_like=np.array(tl), seed=np.array(seed),
                       weights_sha256=np.array(weights_digest(w)), probs=out["probs"], pooled=out["pooled"],
                       embedding_stats=_token_stats(out["embedding_output"]),
                       sequence_stats=_token_stats(out["sequence_output"]), variables=np.array(out["variables"]))
            name = kind.replace("imagebert_", "")
            path = os.path.join(OUT, f"{name}_ref_shim_{tag}.npz")
            if check:
                g = np.load(path)
                for k in ("probs", "pooled", "embedding_stats", "sequence_stats"):
                    worst = max(worst, float(np.abs(g[k] - rec[k]).max()))
                assert list(g["variables"]) == out["variables"]
            else:
                np.savez_compressed(path, **rec)
            print(f"{name}_ref_shim_{tag}: probs[:,1] = {out['probs'][:, 1]}, {len(out['variables'])} variables")
    if check:
        print(f"committed fixtures reproduce: max |diff| = {worst:.2e}")
        assert worst < 1e-5


# ---------------------------------------------------------------------------------------------- record decode KAT
def records_kat():
    """The reference's OWN per-line loader — read_line, seq_padding, and the label-phrase cleaning loop of
    code/imagebert_zk/load_data_v4.py (133-163, 74-85, 34-38), extracted by name (importing the module would read the
    competition data at import time) and run with the reference's own tokenizer class — on synthetic TSV lines."""
    import ast
    import base64
    import importlib.util
    import tempfile
    import types
    rng = np.random.default_rng(11)
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + list("abcdefghijklmnopqrstuvwxyz0123456789'-") + \
            ["women", "men", "leather", "shoes", "dress", "kids", "wash", "basin", "forest", "style", "sen", "department",
             "of", "bag", "hand", "##s", "##ing", "top", "t", "shirt", "others", "red", "black", "连", "衣", "裙"]
    label_lines = ["0\ttop, t-shirt (women)", "1\tleather shoes.", "2\tkids", "3\thand bag, bag", "4\tothers",
                   "5\twash basin", "6\tred black dress for women and men and kids and others"]
    queries = ["women's leather shoes", "sen department of dress", "kids wash basin red", "连衣裙 black", "bag",
               "men t-shirt top forest style others", "red dress", "hand bag for women and men and kids and others red"]
    lines = []
    for i, q in enumerate(queries):
        nb = [1, 2, 3, 12, 2, 1, 3, 2][i]           # 12 > the 10-box budget of the feeds
        h, w = int(rng.integers(200, 900)), int(rng.integers(200, 900))
        boxes = np.round(rng.random((nb, 4)) * 500).astype(np.float32)
        feats = (np.arange(nb * 2048, dtype=np.float32).reshape(nb, 2048) % 17) * 0.125 * (i + 1)   # compressible
        labels = rng.integers(0, 7, nb).astype(np.int64)
        fields = [str(1000 + i), str(h), str(w), str(nb), base64.b64encode(boxes.tobytes()).decode(),
                  base64.b64encode(feats.tobytes()).decode(), base64.b64encode(labels.tobytes()).decode(), q, str(7 * i)]
        lines.append("\t".join(fields) + "\n")
    tf = types.ModuleType("tensorflow")
    tf.gfile = types.SimpleNamespace(GFile=lambda p, m="r": open(p, m, encoding="utf-8"))
    sys.modules["tensorflow"] = tf
    spec = importlib.util.spec_from_file_location("ref_tok_zk2", os.path.join(REF, "code/imagebert_zk/tokenization.py"))
    tokmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tokmod)
    src = open(os.path.join(REF, "code/imagebert_zk/load_data_v4.py")).read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("read_line", "seq_padding", "seq_padding_2")]
    label_loop = [n for n in tree.body if isinstance(n, ast.For) and "multimodal_labels" in ast.get_source_segment(src, n)]
    assert len(fns) == 3 and len(label_loop) == 1
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "data"))
        os.makedirs(os.path.join(d, "work"))
        with open(os.path.join(d, "data", "multimodal_labels.txt"), "w", encoding="utf-8") as f:
            f.write("\n".join(label_lines) + "\n")
        vf = os.path.join(d, "vocab.txt")
        with open(vf, "w", encoding="utf-8") as f:
            f.write("\n".join(vocab) + "\n")
        ns = {"np": np, "base64": base64, "dict_multimodal_labels": {}, "FLAGS": types.SimpleNamespace(sen2forest=0),
              "tokenizer": tokmod.FullTokenizer(vocab_file=vf, do_lower_case=True)}
        os.chdir(os.path.join(d, "work"))
        try:
            exec(compile(ast.Module(body=label_loop + fns, type_ignores=[]), "load_data_v4.py", "exec"), ns)
        finally:
            os.chdir(cwd)
    out = {"lines": np.array(lines), "vocab": np.array(vocab), "label_lines": np.array(label_lines)}
    for flag in (0, 1):                             # evaluate_normal.py / evaluate_normal_sen2fs.py
        ns["FLAGS"].sen2forest = flag
        rec = [ns["read_line"](ln) for ln in lines]
        for i, r in enumerate(rec):
            pid, h, w, nb, boxes5, feats, idx_labels, len_labels, idx_query, qid, query, str_labels = r
            pre = f"s2f{flag}_{i}_"
            out[pre + "scalars"] = np.array([pid, h, w, nb, qid], np.int64)
            out[pre + "boxes5"] = boxes5
            out[pre + "feats_u32"] = np.ascontiguousarray(feats).view(np.uint32)
            out[pre + "label_ids"] = np.asarray(idx_labels, np.int64)
            out[pre + "label_lens"] = np.asarray(len_labels, np.int64)
            out[pre + "query_ids"] = np.asarray(idx_query, np.int64)
            out[pre + "query"] = np.array(query)
    # the batch padding of the feeds (load_data_v4.py:380-389): seq_padding_2 to 10 boxes, zero padding
    rec = [ns["read_line"](ln) for ln in lines]
    out["feats_padded_u32"] = np.ascontiguousarray(ns["seq_padding_2"]([r[5] for r in rec], 10, 0).astype(np.float32)).view(np.uint32)
    out["boxes5_padded"] = ns["seq_padding_2"]([r[4] for r in rec], 10, 0).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "records_kat.npz"), **out)
    print("records_kat:", len(lines), "lines,", os.path.getsize(os.path.join(OUT, "records_kat.npz")), "bytes")


# ---------------------------------------------------------------------------------------------- nDCG@k by the reference's evaluation.py
def ndcg_ref_cases():
    """The reference's OWN evaluate(..., 'ndcg', valid_answer, k) (code/imagebert_lds/src/evaluation.py:4-38, imported as
    is; numpy >= 2 dropped np.asfarray, which is given back as np.asarray(float)) on seeded synthetic rankings."""
    import importlib.util
    import json
    import tempfile
    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a, dtype=float: np.asarray(a, dtype=dtype)
    spec = importlib.util.spec_from_file_location("ref_evaluation", os.path.join(REF, "code/imagebert_lds/src/evaluation.py"))
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    rng = np.random.default_rng(21)
    cases = []
    for ci, (nq, ncand, k) in enumerate([(7, 30, 5), (5, 3, 5), (9, 12, 3), (4, 40, 10)]):
        pred, ans = {}, {}
        for q in range(nq):
            pids = rng.permutation(1000)[:ncand]
            scores = np.round(rng.random(ncand), 3)            # ties on purpose: list.sort is stable
            pred[str(q)] = [[str(int(p)), float(s)] for p, s in zip(pids, scores)]
            n_gt = int(rng.integers(1, 8))
            gt = list(rng.choice(pids, size=min(n_gt, ncand), replace=False)) + ([1234567] if q % 3 == 0 else [])
            ans[str(q)] = [int(x) for x in gt]
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(ans, f)
        val = ev.evaluate(None, None, {q: [list(x) for x in v] for q, v in pred.items()}, "ndcg", f.name, k)
        os.unlink(f.name)
        cases.append({"k": k, "pred": pred, "answers": ans, "ndcg": float(val)})
        print(f"ndcg_ref case {ci}: k={k} ndcg={val:.6f}")
    json.dump({"source": "code/imagebert_lds/src/evaluation.py evaluate(..., 'ndcg', ...)", "cases": cases},
              open(os.path.join(OUT, "ndcg_ref_cases.json"), "w"))


if __name__ == "__main__" and "--ndcg" in sys.argv:
    os.makedirs(OUT, exist_ok=True)
    ndcg_ref_cases()
    sys.exit(0)

def records_kat_lxmert():
    """The same lines through the LXMERT tree's OWN loader functions — read_line, seq_padding, seq_padding_2,
    random_word (code/lxmert/src/utils.py:23-59, 61-96, 126-156; extracted by name: the module imports param, which
    parses argv) — with the LXMERT tree's own BertTokenizer (100-character word limit)."""
    import ast
    import random
    import tempfile
    import types
    g = np.load(os.path.join(OUT, "records_kat.npz"))
    lines = [str(x) for x in g["lines"]]
    label_map = {}
    for ln in g["label_lines"]:
        a = str(ln).split("\t")
        label_map[a[0]] = a[1].replace(",", " ").replace(".", " ").replace("(", " ").replace(")", " ").strip()
    sys.path.insert(0, os.path.join(REF, "code/lxmert/src"))
    import importlib
    lx = importlib.import_module("lxrt.tokenization")
    src = open(os.path.join(REF, "code/lxmert/src/utils.py")).read()
    tree = ast.parse(src)
    want = ("read_line", "seq_padding", "seq_padding_2", "random_word", "random_query", "random_img")
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert len(fns) == len(want)
    import base64
    ns = {"np": np, "base64": base64, "random": random, "MAX_LABLETEXT_LENGTH": 8, "SHUFFLE_RATIO": 0.5,
          "args": types.SimpleNamespace(shuffle_img=False, shuffle_query=False)}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "utils.py", "exec"), ns)
    with tempfile.TemporaryDirectory() as d:
        vf = os.path.join(d, "vocab.txt")
        with open(vf, "w", encoding="utf-8") as f:
            f.write("\n".join(str(t) for t in g["vocab"]) + "\n")
        tok = lx.BertTokenizer(vf, do_lower_case=True)
    random.seed(0)
    out = {}
    rec = [ns["read_line"](ln, label_map, tok) for ln in lines]
    for i, r in enumerate(rec):
        pid, boxes, feats, idx_labels, idx_labels_mask, idx_query, qid = r[:7]
        out[f"{i}_boxes4_f64"] = np.asarray(boxes, np.float64)
        out[f"{i}_label_ids"] = np.asarray(idx_labels, np.int64)
        out[f"{i}_label_mask"] = np.asarray(idx_labels_mask, np.int64)
        out[f"{i}_query_ids"] = np.asarray(idx_query, np.int64)
        out[f"{i}_ids"] = np.array([pid, qid], np.int64)
    pad_boxes, box_mask = ns["seq_padding_2"]([r[1] for r in rec], 10, 0)
    out["boxes4_padded_f32"] = pad_boxes.astype(np.float32)
    out["visn_mask"] = box_mask.astype(np.int64)
    q_pad, q_mask = ns["seq_padding"]([r[5] for r in rec], 23, 0)
    out["query_ids_padded"], out["query_mask"] = q_pad.astype(np.int64), q_mask.astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "records_kat_lxmert.npz"), **out)
    print("records_kat_lxmert:", len(rec), "lines")


if __name__ == "__main__" and "--records" in sys.argv:
    os.makedirs(OUT, exist_ok=True)
    records_kat()
    records_kat_lxmert()
    sys.exit(0)

if __name__ == "__main__" and "--tf-shim" in sys.argv:
    os.makedirs(OUT, exist_ok=True)
    tf_shim_refs(check="--check" in sys.argv)
    sys.exit(0)

if __name__ == "__main__" and "--tokenizer" not in sys.argv:
    os.makedirs(OUT, exist_ok=True)
    ensemble_kat()
    ndcg_kat()
    lxmert_ref()


# ---------------------------------------------------------------------------------------------- tokenizer KAT
def make_tokenizer_golden(out_path=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests",
                                                "golden", "tokenizer_kat.json")):
    """Runs the reference's OWN tokenizer classes (imagebert_zk/tokenization.py with `tensorflow` stubbed: it only uses
    tf.gfile to read the vocab; lxmert/src/lxrt/tokenization.py as is) on a synthetic vocabulary and a set of strings,
    and stores vocabulary, strings, tokens and ids."""
    import importlib
    import importlib.util
    import json
    import tempfile
    import types
    specials = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    singles = list("abcdefghijklmnopqrstuvwxyz0123456789") + list("!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~") + ["α", "β", "é"]
    pieces = ["sen", "department", "of", "forest", "style", "women", "men", "leather", "shoes", "shoe", "wash", "basin",
              "run", "##ning", "##ing", "##s", "##ed", "un", "##aff", "##able", "t", "shirt", "kids", "cafe", "naive",
              "hello", "world", "2020", "##20", "20", "3d", "x", "##x", "##xx", "##xxx", "xxxx", "zero", "tab", "sep",
              "女", "士", "皮", "鞋", "包", "##a", "##b", "##c", "the", "for", "and", "black", "white", "red", "dress"]
    vocab = specials + singles + pieces
    texts = ["sen department of Women's leather shoes", "forest style  women dress", "女士皮鞋 black 包", "Running shoes for MEN",
             "unaffable T-shirt (2020)", "café naïve HELLO wörld", "kids' wash\tbasin red", "​zero emoji\U0001F600 end",
             "x" * 120, "x" * 250, "α-β ３Ｄ", "###ing ##s", "", "   ", "unknownword qqq", "shoes,shoe.shoes!", "ＡＢＣ abc",
             "the red-and-white dress for women and men", "20202020 20 2020a"]
    tf = types.ModuleType("tensorflow")
    tf.gfile = types.SimpleNamespace(GFile=lambda p, m="r": open(p, m, encoding="utf-8"))
    sys.modules.setdefault("tensorflow", tf)
    with tempfile.TemporaryDirectory() as d:
        vf = os.path.join(d, "vocab.txt")
        with open(vf, "w", encoding="utf-8") as f:
            f.write("\n".join(vocab) + "\n")
        spec = importlib.util.spec_from_file_location("ref_tok_zk", os.path.join(REF, "code/imagebert_zk/tokenization.py"))
        zk = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(zk)
        sys.path.insert(0, os.path.join(REF, "code/lxmert/src"))
        lx = importlib.import_module("lxrt.tokenization")
        tz = zk.FullTokenizer(vocab_file=vf, do_lower_case=True)
        tl = lx.BertTokenizer(vf, do_lower_case=True)
        cases = []
        for t in texts:
            a, b = tz.tokenize(t), tl.tokenize(t)
            cases.append({"text": t, "tf_tokens": a, "tf_ids": tz.convert_tokens_to_ids(a), "lxmert_tokens": b,
                          "lxmert_ids": tl.convert_tokens_to_ids(b)})
    with open(out_path, "w", encoding="utf-8") as f:
        json.dump({"vocab": vocab, "cases": cases}, f, ensure_ascii=False, indent=0)
    print("wrote", out_path, len(cases), "cases")


if __name__ == "__main__" and "--tokenizer" in sys.argv:
    make_tokenizer_golden()
