"""End-to-end scoring drivers: competition TSV -> decode -> feeds -> scorer -> the score file of each model -> ensemble.

These are the loops around the hot path that the reference spreads over three scripts and `code/main.py`:

  imagebert_zk/evaluate_normal.py:210-252         ImageBertB / C: restore the EMA checkpoint, `next(data_generator)`,
                                                  sess.run, one `qid \\t pid \\t prob[1]` line per pair
  imagebert_zk/evaluate_normal_sen2fs.py          the same model on the "sen department of" -> "forest style" rewrite
  imagebert_lds/src/run_pretraining_predict_score.py:558-589   ImageBertA: 29005 / 5 batches, `qid \\t pid \\t prob[1]`
  lxmert/src/tasks/kdd_model.py:46-129            KDD.predict: Softmax(1)(logit)[:, -1], CSV `query-id,product-id,score`
  code/main.py:11-104                             the 4-file ensemble -> submission.csv

Here one function (`score_tsv`) does the per-model loop for all three scorers: lines are decoded 256 at a time by the
C++ decoder straight into pinned arrays (three decoders in rotation, the decode of chunk i + 1 running on a helper
thread while chunk i is assembled, copied and scored), feeds are assembled (cached WordPiece ids of queries and label phrases, box normalisation on the
GPU) and go through MatchScorer.score_stream; with `world > 1` every rank decodes and scores only its contiguous range
of lines and the scores meet in one all-gather.  `run_ensemble` chains the three models and `ensemble.main`.
Errors are raised, not swallowed (the reference ends its loop on a bare `except:`, evaluate_normal.py:250).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ensemble, records
from .config import LXMERT, ZK
from .scorer import MatchScorer, shard_range, sharded_score_stream


def read_tsv_lines(path: str) -> List[bytes]:
    """Data lines of a competition TSV; header lines are skipped as the loaders do (`if "product_id" in line`,
    load_data_v4.py:222, 237)."""
    out = []
    with open(path, "rb") as f:
        for line in f:
            if b"product_id" in line or not line.strip():
                continue
            out.append(line)
    return out


def _ids_of(lines: Sequence[bytes]):
    """(query_id, product_id) of every line without splitting the 300 KB base64 fields in between."""
    q = np.empty(len(lines), np.int64)
    p = np.empty(len(lines), np.int64)
    for i, line in enumerate(lines):
        p[i] = int(line[:line.index(b"\t")])
        q[i] = int(line[line.rindex(b"\t") + 1:])          # (int() ignores the line end; no copy of the 300 KB line)
    return q, p


def score_tsv(scorer: MatchScorer, tokenizer, label_map: Dict[int, str], lines: Sequence[bytes],
              sen2forest: bool = False, rank: int = 0, world: int = 1, n_threads: int = 0) -> Dict[str, np.ndarray]:
    """Scores every TSV record of `lines` with `scorer`; returns query_id [N], product_id [N], score [N] (fp32, the
    probability of the positive class: probs[:, 1] / Softmax(1)(logit)[:, -1]) in file order, on every rank."""
    cfg = scorer.cfg
    asm = records.FeedAssembler(cfg, tokenizer, label_map, sen2forest=sen2forest)
    # Three decoders in rotation: while chunk k is assembled, copied and scored, a helper thread already decodes chunk
    # k + 1 (the decode is one GIL-free C call) into the arrays chunk k - 2 used, whose copies score_stream has waited
    # for before it asked for chunk k.
    decoders = [records.RecordDecoder(scorer.max_batch, max_boxes=cfg.nbox, feat_dim=cfg.feat_dim, n_threads=n_threads)
                for _ in range(3)]
    shard_hi = shard_range(len(lines), rank, world)[1]
    helper = ThreadPoolExecutor(max_workers=1)
    calls, ahead = [0], {}

    def fetch(lo, hi):
        k = calls[0]
        calls[0] += 1
        fut = ahead.pop((lo, hi), None)
        batch = fut.result() if fut is not None else decoders[k % 3].decode(lines[lo:hi])
        nlo, nhi = hi, min(hi + scorer.max_batch, shard_hi)
        if nlo < nhi:
            ahead[(nlo, nhi)] = helper.submit(decoders[(k + 1) % 3].decode, lines[nlo:nhi])
        with torch.cuda.stream(scorer.copy_stream):      # the box normalisation kernel runs where the slot copies do
            return asm.assemble(batch, device=scorer.device)

    try:
        scores = sharded_score_stream(scorer, len(lines), fetch, rank, world)
    finally:
        helper.shutdown(wait=True)
    q, p = _ids_of(lines)
    return {"query_id": q, "product_id": p, "score": scores.numpy().astype(np.float32)}


def write_scores(path: str, result: Dict[str, np.ndarray], lxmert_csv: bool = False) -> None:
    """The per-model score file in the reference's format (see ensemble.write_score_file)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    ensemble.write_score_file(path, result["query_id"].tolist(), result["product_id"].tolist(), result["score"].tolist(),
                              lxmert_csv=lxmert_csv)


def run_ensemble(scorers: Dict[str, MatchScorer], tokenizers: Dict[str, object], label_map: Dict[int, str],
                 tsv_path: str, out_dir: str, rank: int = 0, world: int = 1,
                 lds_key: str = "imagebert_lds") -> List:
    """The whole prediction flow of the reference's README: the four score files (ImageBertB, ImageBertB on the
    sen2forest rewrite, ImageBertA, LXMERT) under `out_dir` with the reference's file names, then code/main.py.
    Returns the submission rows; rank 0 writes the files."""
    lines = read_tsv_lines(tsv_path)
    names = {"zk": "testB_result_match_keyword_valid_finetune_251.txt",
             "zk_s2f": "testB_result_match_keyword_valid_finetune_251_sen_to_forest.txt",
             "lds": "testBscore_imagebert.txt", "lxmert": "testB_score_lxmert.csv"}
    res = {
        "zk": score_tsv(scorers[ZK], tokenizers[ZK], label_map, lines, False, rank, world),
        "zk_s2f": score_tsv(scorers[ZK], tokenizers[ZK], label_map, lines, True, rank, world),
        "lds": score_tsv(scorers[lds_key], tokenizers[lds_key], label_map, lines, False, rank, world),
        "lxmert": score_tsv(scorers[LXMERT], tokenizers[LXMERT], label_map, lines, False, rank, world),
    }
    paths = {k: os.path.join(out_dir, n) for k, n in names.items()}
    if rank == 0:
        for k in res:
            write_scores(paths[k], res[k], lxmert_csv=(k == "lxmert"))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return ensemble.main(paths["zk"], paths["zk_s2f"], paths["lds"], paths["lxmert"],
                         os.path.join(out_dir, "submission.csv") if rank == 0 else os.devnull)
