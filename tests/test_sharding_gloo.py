"""N > 1 host logic on CPU: two gloo ranks shard a pair list, "score" their ranges with a stand-in scorer (the real one
needs a B200), all-gather, and every rank must end up with the full score vector in the original pair order -- then
run the ensemble on it.  What is exercised is exactly what runs around the kernels on a multi-GPU box:
scorer.shard_range / sharded_score / allgather_scores (one all_gather_into_tensor, the only collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ensemble
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import (allgather_scores, distinct_queries, shard_range,
                                                                      sharded_score, sharded_score_stream)


class _StandInScorer:
    """Same surface as MatchScorer.score / .spec / .device; the 'score' is a deterministic function of the feeds."""
    spec = {"query_ids": None, "feats": None}
    device = torch.device("cpu")

    def score(self, feeds):
        # integer arithmetic: bit-identical whatever the chunking (as the real kernels are, test_baseline_size_properties)
        s = ((feeds["query_ids"].long().sum(1) * 31 + (feeds["feats"] > 0).long().sum(dim=(1, 2)) * 7) % 1000).float() / 1000.0
        return torch.stack([1.0 - s, s], 1)

    def score_stream(self, n, fetch):
        """MatchScorer.score_stream's contract: fetch(lo, hi) in LOCAL indices, chunks of at most max_batch."""
        out = [self.score(fetch(lo, min(n, lo + 5))) for lo in range(0, n, 5)]
        return torch.cat(out) if out else torch.empty((0, 2))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _feeds(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"query_ids": torch.randint(0, 1000, (n, 20), generator=g, dtype=torch.int32),
            "feats": torch.randn(n, 4, 16, generator=g),
            "not_a_feed": torch.zeros(n)}          # extra keys are ignored by sharded_score


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        feeds = _feeds(n)
        full = sharded_score(_StandInScorer(), feeds, rank, world)
        want = _StandInScorer().score(feeds)[:, 1]
        assert full.shape == (n,), full.shape
        assert torch.equal(full, want), (rank, (full - want).abs().max())
        # the streamed variant: every rank fetches only pairs of its own range, in global indices
        seen = []

        def fetch(lo, hi):
            seen.append((lo, hi))
            return {k: v[lo:hi] for k, v in feeds.items()}
        streamed = sharded_score_stream(_StandInScorer(), n, fetch, rank, world)
        assert torch.equal(streamed, want)
        lo, hi, _ = shard_range(n, rank, world)
        assert all(lo <= a and b <= hi for a, b in seen) and sum(b - a for a, b in seen) == hi - lo
        # ragged: rank r contributes r+1 real scores, padded to 3
        mine = torch.full((3,), -1.0)
        mine[: rank + 1] = float(rank + 1)
        got = allgather_scores(mine, 3 * world, world)
        assert got.tolist() == sum(([float(r + 1)] * (r + 1) + [-1.0] * (2 - r) for r in range(world)), [])
        # the ensemble runs on the gathered vector on every rank and must agree across ranks
        qids = [f"q{i // 6}" for i in range(n)]
        pids = [f"p{i}" for i in range(n)]
        d = ensemble.scores_from_arrays(qids, pids, full.tolist())
        rows, _ = ensemble.merge_and_select(d, d, d, d)
        np.save(os.path.join(out_dir, f"rows_{rank}.npy"), np.array([r[1] for r in rows], dtype=object),
                allow_pickle=True)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 64, 3])
def test_two_gloo_ranks_shard_gather_and_ensemble(n, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rows_0.npy", allow_pickle=True)
    r1 = np.load(tmp_path / "rows_1.npy", allow_pickle=True)
    assert len(r0) == len(r1) and all(list(a) == list(b) for a, b in zip(r0, r1))


class _StandInTsvScorer:
    """What drivers.score_tsv touches of a MatchScorer: cfg, max_batch, device, copy_stream, score_stream."""
    device = torch.device("cpu")
    copy_stream = None                      # torch.cuda.stream(None) is a no-op
    max_batch = 4

    def __init__(self, cfg):
        self.cfg = cfg
        self.fetched = []

    def score_stream(self, n, fetch, out=None, on_device=False):
        outs = []
        for lo in range(0, n, self.max_batch):
            f = fetch(lo, min(n, lo + self.max_batch))
            self.fetched.append(int(f["query_ids"].shape[0]))
            # integers only: independent of the chunking; depends on every decoded array and on the query text
            s = (f["feats"].double().mul(64).round().long().sum(dim=(1, 2)) + f["label_ids"].long().sum(dim=(1, 2)) * 3
                 + f["query_ids"].long().sum(1) * 7) % 1000
            outs.append(torch.stack([1.0 - s.float() / 1000.0, s.float() / 1000.0], 1))
        return torch.cat(outs) if outs else torch.empty((0, 2))


def _tsv_worker(rank, world, port, n, out_dir):
    """drivers.score_tsv on two gloo ranks: every rank decodes only its own lines (one chunk ahead, on the helper
    thread, three decoders in rotation) and all ranks end with the scores of the whole file in file order."""
    import json

    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import drivers, tokenizer
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LDS, ModelConfig
    from tests.test_widening_host import _make_lines
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tokenizer_kat.json"), encoding="utf-8"))
        vocab = {t: i for i, t in enumerate(dict.fromkeys(kat["vocab"]))}
        tok = tokenizer.FullTokenizer(vocab=vocab)
        labels = {i: "women dress" for i in range(33)}
        lines, _ = _make_lines(n, np.random.default_rng(5), max_nb=12)
        cfg = ModelConfig(LDS, n_layers=1, lq=20, nbox=10, vocab=len(vocab))
        whole = drivers.score_tsv(_StandInTsvScorer(cfg), tok, labels, lines)          # one rank, the whole file
        sc = _StandInTsvScorer(cfg)
        got = drivers.score_tsv(sc, tok, labels, lines, rank=rank, world=world)
        lo, hi, _ = shard_range(n, rank, world)
        assert sum(sc.fetched) == hi - lo and all(0 < c <= sc.max_batch for c in sc.fetched)
        assert np.array_equal(got["score"], whole["score"]) and len(got["score"]) == n
        assert got["query_id"].tolist() == [7 * i for i in range(n)] and got["product_id"].tolist() == [1000 + i for i in range(n)]
        np.save(os.path.join(out_dir, f"tsv_{rank}.npy"), got["score"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [23, 8, 1])
def test_two_gloo_ranks_score_a_tsv_through_the_driver(n, tmp_path):
    mp.spawn(_tsv_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    assert np.array_equal(np.load(tmp_path / "tsv_0.npy"), np.load(tmp_path / "tsv_1.npy"))


def test_shard_range_covers_every_pair_once():
    for n in (0, 1, 7, 30000, 10001):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_distinct_queries_groups_pairs_by_query_and_mask():
    """The host side of mmr_inputs.lang_unique / lang_slot: representatives ascending, every pair mapped to the first
    pair with the same (ids, mask) row; two queries with equal ids but different masks stay apart."""
    q = np.array([[5, 6, 0], [1, 2, 0], [5, 6, 0], [1, 2, 0], [9, 9, 9], [5, 6, 0]])
    m = np.array([[1, 1, 0], [1, 1, 0], [1, 1, 0], [1, 1, 1], [1, 1, 1], [1, 1, 0]])
    uniq, slot = distinct_queries(q, m)
    assert uniq.dtype == torch.int32 and slot.dtype == torch.int32
    assert uniq.tolist() == [0, 1, 3, 4] and slot.tolist() == [0, 1, 0, 2, 3, 0]
    rng = np.random.default_rng(0)
    pool = rng.integers(0, 50, (7, 12))
    owner = rng.integers(0, 7, 200)
    uniq, slot = distinct_queries(pool[owner], np.ones((200, 12), np.int64))
    u, s_ = uniq.numpy(), slot.numpy()
    assert (np.diff(u) > 0).all() and (pool[owner][u[s_]] == pool[owner]).all() and (u[s_] <= np.arange(200)).all()
    assert len(u) == len(np.unique(owner))
