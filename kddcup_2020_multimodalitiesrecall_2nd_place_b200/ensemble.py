"""Ensemble scorer of code/main.py: weighted 4-way merge, product-uniqueness filter, per-query top-5.

Same file contract as the reference (main.py:11-39, 88-104): three TSV files `qid \\t pid \\t score` (ImageBertB,
ImageBertB with the sen2forest rewrite, ImageBertA), one CSV with a header line containing "query" (LXMERT), output
CSV with header `query-id,product1..product5`.  Same semantics, including the quirks: scores missing from a TSV
file are back-filled from LXMERT (main.py:50-58), iteration runs over LXMERT's product set of each query of the
first file, a product survives only if its best merged score beats its second best by >= 0.92 (strict `< 0.92`
skip, main.py:80-82) and the pair holds that best score within 1e-5 (main.py:83), queries with 1..4 survivors fall
back to the unfiltered top-5 (main.py:101-104) and queries with no survivor at all are not written.

The arithmetic is vectorised numpy in fp64 (the reference computes in Python floats = fp64) over flat arrays in
LXMERT file order; it is O(N log N) host work on <= 30 k scores (negligible next to scoring),
so it stays on the host — scores arrive there anyway through the all-gather of scorer.allgather_scores.
"""
from __future__ import annotations

import csv
from collections import OrderedDict
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

WEIGHTS = (0.2, 0.2, 0.3, 0.3)   # main.py:59 — zk, zk_sen2forest, lds (ImageBertA), lxmert
MARGIN = 0.92                    # main.py:81
TIE = 1e-5                       # main.py:83
TOPK = 5

ScoreDict = Dict[str, Dict[str, float]]


def read_scores(path: str, sep: str = "\t", header_token: str | None = None) -> ScoreDict:
    """main.py:11-39: later duplicates of a (qid, pid) overwrite earlier ones; insertion order = file order."""
    out: ScoreDict = OrderedDict()
    with open(path) as f:
        for line in f:
            if header_token is not None and header_token in line:
                continue
            a = line.strip().split(sep)
            out.setdefault(a[0], OrderedDict())[a[1]] = float(a[2])
    return out


def scores_from_arrays(qids: Iterable, pids: Iterable, scores: Iterable) -> ScoreDict:
    out: ScoreDict = OrderedDict()
    for q, p, s in zip(qids, pids, scores):
        out.setdefault(str(q), OrderedDict())[str(p)] = float(s)
    return out


def _flatten(d1: ScoreDict, d2: ScoreDict, d3: ScoreDict, d4: ScoreDict):
    """Flat (query index, product index, 4 score columns) in the reference's iteration order, back-filled."""
    q_names: List[str] = []
    p_index: Dict[str, int] = {}
    qi, pi, cols = [], [], ([], [], [], [])
    for qid in d1:                                    # main.py:44 (KeyError if another file lacks the query)
        r1, r2, r3, r4 = d1[qid], d2[qid], d3[qid], d4[qid]
        q = len(q_names)
        q_names.append(qid)
        for pid, s4 in r4.items():                    # main.py:49
            qi.append(q)
            pi.append(p_index.setdefault(pid, len(p_index)))
            cols[0].append(r1.get(pid, s4))           # main.py:50-58 back-fill
            cols[1].append(r2.get(pid, s4))
            cols[2].append(r3.get(pid, s4))
            cols[3].append(s4)
    p_names = [None] * len(p_index)
    for name, i in p_index.items():
        p_names[i] = name
    S = np.array(cols, dtype=np.float64)
    return q_names, p_names, np.array(qi, np.int64), np.array(pi, np.int64), S


def select_flat(qi: np.ndarray, pi: np.ndarray, S: np.ndarray, n_products: int,
                weights: Sequence[float] = WEIGHTS, margin: float = MARGIN, tie: float = TIE, topk: int = TOPK):
    """The arithmetic of main.py:59-104 on flat arrays: qi / pi = query / product index of every pair (pairs of a query
    contiguous, in the reference's iteration order), S [4, n] fp64 score columns.  Returns (rows, merged): rows =
    [(first pair index of the query, [pair index] * <= topk)] in the reference's output order (filtered queries in
    first-seen order, then the < topk fall-backs), merged = fp64 merged score per pair."""
    n = qi.shape[0]
    if n == 0:
        return [], np.zeros(0)
    w1, w2, w3, w4 = weights
    merged = w1 * S[0] + w2 * S[1] + w3 * S[2] + w4 * S[3]      # main.py:59, same left-to-right fp64 evaluation

    # best and runner-up merged score of every product over all queries (main.py:65-72, 78-82)
    order = np.lexsort((-merged, pi))                           # by product, then descending score
    sp = pi[order]
    first = np.r_[True, sp[1:] != sp[:-1]]
    starts = np.nonzero(first)[0]
    counts = np.diff(np.r_[starts, n])
    best = np.empty(n_products)
    second = np.full(n_products, -np.inf)
    best[sp[starts]] = merged[order[starts]]
    multi = counts >= 2
    second[sp[starts[multi]]] = merged[order[starts[multi] + 1]]
    unique_enough = ~((best - second) < margin)                  # skip iff a[0] - a[1] < 0.92; single score survives
    keep = unique_enough[pi] & (np.abs(merged - best[pi]) < tie)  # main.py:83

    rows, short = [], []
    # per-query segments are contiguous in the flattened order
    q_starts = np.nonzero(np.r_[True, qi[1:] != qi[:-1]])[0]
    q_ends = np.r_[q_starts[1:], n]
    for a, b in zip(q_starts, q_ends):
        idx = np.arange(a, b)
        surv = idx[keep[a:b]]
        if surv.size == 0:
            continue                                            # never enters dict_eval_merge_top1: not written
        if surv.size < topk:
            short.append(int(a))
            continue
        rows.append((int(a), surv[np.argsort(-merged[surv], kind="stable")[:topk]]))   # sorted(reverse=True) is stable
    seg_end = dict(zip(q_starts.tolist(), q_ends.tolist()))
    for a in short:                                             # main.py:101-104
        idx = np.arange(a, seg_end[a])
        rows.append((a, idx[np.argsort(-merged[idx], kind="stable")[:topk]]))
    return rows, merged


def merge_and_select(d1: ScoreDict, d2: ScoreDict, d3: ScoreDict, d4: ScoreDict,
                     weights: Sequence[float] = WEIGHTS, margin: float = MARGIN, tie: float = TIE,
                     topk: int = TOPK) -> Tuple[List[Tuple[str, List[str]]], np.ndarray]:
    """Returns (rows, merged): rows = [(qid, [pid] * topk)] in the reference's output order (filtered queries in
    first-seen order, then the < topk fall-backs), merged = fp64 merged score per flattened pair."""
    q_names, p_names, qi, pi, S = _flatten(d1, d2, d3, d4)
    flat_rows, merged = select_flat(qi, pi, S, len(p_names), weights, margin, tie, topk)
    rows = [(q_names[qi[a]], [p_names[pi[i]] for i in top]) for a, top in flat_rows]
    return rows, merged


def merge_and_select_gpu(d1: ScoreDict, d2: ScoreDict, d3: ScoreDict, d4: ScoreDict,
                         weights: Sequence[float] = WEIGHTS, margin: float = MARGIN, tie: float = TIE,
                         topk: int = TOPK, device: int = 0) -> Tuple[List[Tuple[str, List[str]]], np.ndarray]:
    """Same result as merge_and_select, computed by the device kernels of csrc/ensemble.cu (mmr_ensemble_topk):
    flattening / back-filling of the score files stays on the host (dictionary work), merge + uniqueness filter + top-k
    run on the GPU.  Raises without an sm_100 device (no CPU fallback behind this name)."""
    import ctypes as C

    import torch

    from . import _lib
    lib = _lib.load()
    q_names, p_names, qi, pi, S = _flatten(d1, d2, d3, d4)
    n = int(qi.shape[0])
    if n == 0:
        return [], np.zeros(0)
    dev = torch.device("cuda", device)
    q_starts = np.nonzero(np.r_[True, qi[1:] != qi[:-1]])[0]
    query_start = torch.from_numpy(np.r_[q_starts, n].astype(np.int32)).to(dev)
    nq, P = len(q_starts), len(p_names)
    s = [torch.from_numpy(np.ascontiguousarray(S[k], dtype=np.float64)).to(dev) for k in range(4)]
    product_of = torch.from_numpy(pi.astype(np.int32)).to(dev)
    merged = torch.empty(n, dtype=torch.float64, device=dev)
    top = torch.empty((nq, topk), dtype=torch.int32, device=dev)
    status = torch.empty(nq, dtype=torch.int32, device=dev)
    ws = torch.empty(20 * P, dtype=torch.uint8, device=dev)
    w = (C.c_double * 4)(*[float(x) for x in weights])
    lib.mmr_ensemble_topk.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_double,
                                                        C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                        C.c_void_p, C.c_size_t, C.c_void_p]
    _lib.check(lib.mmr_ensemble_topk(s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), s[3].data_ptr(),
                                     product_of.data_ptr(), query_start.data_ptr(), n, nq, P, w, float(margin),
                                     float(tie), int(topk), merged.data_ptr(), top.data_ptr(), status.data_ptr(),
                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    top_h, status_h = top.cpu().numpy(), status.cpu().numpy()
    rows: List[Tuple[str, List[str]]] = []
    for want in (1, 2):                                         # filtered queries first, then the fall-backs
        for k, a in enumerate(q_starts):
            if status_h[k] == want:
                rows.append((q_names[qi[a]], [p_names[pi[i]] for i in top_h[k] if i >= 0]))
    return rows, merged.cpu().numpy()


def write_submission(path: str, rows: List[Tuple[str, List[str]]]) -> None:
    """main.py:88-90, 99, 104."""
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["query-id", "product1", "product2", "product3", "product4", "product5"])
        for qid, pids in rows:
            w.writerow([qid, *pids])


def write_score_file(path: str, qids, pids, scores, lxmert_csv: bool = False) -> None:
    """Per-model score files as the reference drivers write them: `qid \\t pid \\t score` (evaluate_normal.py:247,
    run_pretraining_predict_score.py:585-589) or `query-id,product-id,score` with header (kdd_model.py:117-128)."""
    with open(path, "w") as f:
        if lxmert_csv:
            f.write("query-id,product-id,score\n")
        sep = "," if lxmert_csv else "\t"
        for q, p, s in zip(qids, pids, scores):
            f.write(f"{q}{sep}{p}{sep}{float(s)!r}\n")


def main(zk="../prediction_result/testB_result_match_keyword_valid_finetune_251.txt",
         zk_s2f="../prediction_result/testB_result_match_keyword_valid_finetune_251_sen_to_forest.txt",
         lds="../prediction_result/testBscore_imagebert.txt",
         lxmert="../prediction_result/testB_score_lxmert.csv",
         out="../prediction_result/submission.csv"):
    """Drop-in for `python2 code/main.py` (default paths are the reference's)."""
    rows, _ = merge_and_select(read_scores(zk), read_scores(zk_s2f), read_scores(lds),
                               read_scores(lxmert, ",", "query"))
    write_submission(out, rows)
    return rows


if __name__ == "__main__":
    import sys
    main(*sys.argv[1:])
