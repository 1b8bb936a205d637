"""CPU tests of the product ensemble (kddcup_..._b200/ensemble.py) against the reference's shipped files and
against the oracle restatement of code/main.py on randomised score sets."""
import os
from collections import OrderedDict

import numpy as np
import pytest

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ensemble as prod
from oracle import ensemble as oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _kat_dicts(module_scores_from_arrays):
    g = np.load(os.path.join(GOLD, "ensemble_kat.npz"))
    ds = [module_scores_from_arrays(g[t + "_q"].tolist(), g[t + "_p"].tolist(), g[t + "_s"].tolist())
          for t in ("zk", "zk_s2f", "lds", "lxmert")]
    want = OrderedDict((str(r[0]), [str(x) for x in r[1:]]) for r in g["submission"].tolist())
    return ds, want


def test_known_answer_submission_csv_exact():
    """The four shipped per-model score files -> prediction_result/submission.csv, 994/994 rows in exact order of
    products (code/main.py:11-104)."""
    ds, want = _kat_dicts(prod.scores_from_arrays)
    rows, merged = prod.merge_and_select(*ds)
    assert len(rows) == 994
    assert dict(rows) == dict(want)
    assert merged.shape[0] == 29005


def _random_case(seed, n_q=40, n_p=120):
    rng = np.random.default_rng(seed)
    files = [[], [], [], []]
    for q in range(n_q):
        k = int(rng.integers(3, 12))
        pids = rng.choice(n_p, size=k, replace=False)
        for p in pids:
            base = rng.random() ** 3 if rng.random() < 0.7 else 0.9 + 0.1 * rng.random()
            for f in range(4):
                if f < 3 and rng.random() < 0.03:
                    continue                       # missing from a TSV file -> back-filled from lxmert
                files[f].append((str(q), str(p), float(np.clip(base + 0.05 * rng.standard_normal(), 0, 1))))
        if rng.random() < 0.2:                     # duplicate line: the later one wins
            q_, p_, _ = files[3][-1]
            for f in range(4):
                files[f].append((q_, p_, float(rng.random())))
    return files


@pytest.mark.parametrize("seed", range(8))
def test_matches_oracle_on_random_scores(seed):
    files = _random_case(seed)
    a = [prod.scores_from_arrays(*zip(*f)) for f in files]
    b = [oracle.OrderedDict() for _ in files]
    for d, f in zip(b, files):
        for q, p, s in f:
            d.setdefault(q, OrderedDict())[p] = s
    rows_p, merged_p = prod.merge_and_select(*a)
    rows_o, merged_o = oracle.merge_and_select(*b)
    assert rows_p == rows_o                                   # same queries, same order, same products
    flat_o = np.array([v for q in merged_o for v in merged_o[q].values()])
    assert np.array_equal(merged_p, flat_o)                   # bit-identical fp64 merge


def test_file_round_trip(tmp_path):
    files = _random_case(99, n_q=6, n_p=20)
    paths = []
    for i, f in enumerate(files):
        p = tmp_path / f"m{i}.txt"
        q, pid, s = zip(*f)
        prod.write_score_file(str(p), q, pid, s, lxmert_csv=(i == 3))
        paths.append(str(p))
    out = tmp_path / "submission.csv"
    rows = prod.main(*paths, out=str(out))
    lines = out.read_text().strip().split("\n")
    assert lines[0] == "query-id,product1,product2,product3,product4,product5"
    assert [ln.split(",") for ln in lines[1:]] == [[q, *p] for q, p in rows]
    d = [prod.scores_from_arrays(*zip(*f)) for f in files]
    assert rows == prod.merge_and_select(*d)[0]


def test_missing_query_raises_like_the_reference():
    d1 = prod.scores_from_arrays(["1"], ["7"], [0.5])
    empty = OrderedDict()
    with pytest.raises(KeyError):
        prod.merge_and_select(d1, d1, d1, empty)
