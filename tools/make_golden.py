"""Generates the committed golden fixtures under tests/golden/ from the reference tree (dev container only).

    python tools/make_golden.py

  ensemble_kat.npz  the four shipped per-model score files + the shipped submission.csv (code/main.py KAT)
  ndcg_kat.npz      shipped valid scores + valid answers (nDCG@5 known answer 0.7098, report table 5)
  lxmert_ref_*.npz  outputs of the reference's OWN LXMERT code (oracle/lxmert_ref.py) on seeded synthetic
                    weights/inputs from kddcup_2020_multimodalitiesrecall_2nd_place_b200.synth
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


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ensemble_kat()
    ndcg_kat()
    lxmert_ref()
