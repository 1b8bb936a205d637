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
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import allgather_scores, shard_range, sharded_score


class _StandInScorer:
    """Same surface as MatchScorer.score / .spec / .device; the 'score' is a deterministic function of the feeds."""
    spec = {"query_ids": None, "feats": None}
    device = torch.device("cpu")

    def score(self, feeds):
        # integer arithmetic: bit-identical whatever the chunking (as the real kernels are, test_baseline_size_properties)
        s = ((feeds["query_ids"].long().sum(1) * 31 + (feeds["feats"] > 0).long().sum(dim=(1, 2)) * 7) % 1000).float() / 1000.0
        return torch.stack([1.0 - s, s], 1)


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


def test_shard_range_covers_every_pair_once():
    for n in (0, 1, 7, 30000, 10001):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(n))
