"""Restatement of code/main.py (ensemble + product-uniqueness filter + top-5) and of nDCG@k.
TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned by the reference's shipped files:
prediction_result/*.txt|csv -> prediction_result/submission.csv (994 rows, exact order) and nDCG@5 = 0.7098.

Iteration order matters (dict insertion order == file order under CPython 3.7+; the reference ran Python 2 where
dict order is arbitrary, but every order-dependent step below is order-independent in its RESULT except the row
order of the CSV, which we emit in first-seen query order and compare as a set of rows keyed by query id).
"""
from __future__ import annotations

import math
from collections import OrderedDict


def read_score_file(path, sep="\t", skip_header_token=None):
    """main.py:11-39."""
    d = OrderedDict()
    for line in open(path):
        if skip_header_token is not None and skip_header_token in line:
            continue
        arr = line.strip().split(sep)
        d.setdefault(arr[0], OrderedDict())[arr[1]] = float(arr[2])
    return d


def merge_and_select(d1, d2, d3, d4, weights=(0.2, 0.2, 0.3, 0.3), margin=0.92, tie=1e-5, topk=5):
    """main.py:41-104.  d1..d4: {qid: {pid: score}} for zk, zk-sen2forest, lds (ImageBertA), lxmert.
    Returns (rows, merged) with rows = [(qid, [pid x topk])] and merged = {qid: {pid: score}}."""
    w1, w2, w3, w4 = weights
    best = {}
    all_scores = {}
    merged = OrderedDict()
    for qid in d1:                                             # :44
        r1, r2, r3, r4 = d1[qid], d2[qid], d3[qid], d4[qid]
        for pid in r4:                                         # :49 iterate over lxmert's product set
            if pid not in r1:
                r1[pid] = r4[pid]                              # :50-58 back-fill from lxmert
            if pid not in r2:
                r2[pid] = r4[pid]
            if pid not in r3:
                r3[pid] = r4[pid]
            m = w1 * r1[pid] + w2 * r2[pid] + w3 * r3[pid] + w4 * r4[pid]   # :59
            merged.setdefault(qid, OrderedDict())[pid] = m
            if pid not in best or m > best[pid]:
                best[pid] = m                                  # :65-68
            all_scores.setdefault(pid, []).append(m)           # :69-72
    top1 = OrderedDict()
    for qid in merged:                                         # :76-86
        for pid in merged[qid]:
            a = sorted(all_scores[pid], reverse=True)
            if len(a) >= 2 and a[0] - a[1] < margin:
                continue
            if abs(merged[qid][pid] - best[pid]) < tie:
                top1.setdefault(qid, OrderedDict())[pid] = merged[qid][pid]
    rows, short = [], []
    for qid in top1:                                           # :92-99
        s = sorted(top1[qid].items(), key=lambda kv: kv[1], reverse=True)
        if len(s) < topk:
            short.append(qid)
            continue
        rows.append((qid, [p for p, _ in s[:topk]]))
    for qid in short:                                          # :101-104 fall back to the unfiltered merge
        s = sorted(merged[qid].items(), key=lambda kv: kv[1], reverse=True)
        rows.append((qid, [p for p, _ in s[:topk]]))
    return rows, merged


def dcg_at_k(r, k):
    """imagebert_lds/src/evaluation.py dcg_at_k: sum r_i / log2(i+2)."""
    r = list(r)[:k]
    return sum(v / math.log2(i + 2) for i, v in enumerate(r))


def ndcg_at_k(pred_by_query, answers, k=5):
    """evaluation.py:4-38 / evaluate_function.py:5-45: mean over queries of DCG@k / ideal DCG@k where relevance is
    membership in the ground-truth list and the ideal has min(k, |gt|) ones."""
    total, n = 0.0, 0
    for qid, gt in answers.items():
        if qid not in pred_by_query:
            continue
        preds = pred_by_query[qid][:k]
        rel = [1.0 if p in gt else 0.0 for p in preds]
        ideal = [1.0] * min(k, len(gt))
        idcg = dcg_at_k(ideal, k)
        total += dcg_at_k(rel, k) / idcg if idcg > 0 else 0.0
        n += 1
    return total / max(n, 1)
