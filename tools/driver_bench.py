"""TSV -> scores throughput of the end-to-end driver (drivers.score_tsv): C++ record decode + feed assembly + H2D +
the 12-layer forward + scores back, on synthetic competition-shaped lines held in memory (the file read is not timed).

    python tools/driver_bench.py [n_lines] [model] [--host-only]

Shapes: BASELINE configs[1] (32 query tokens, 36 region slots, 12 layers); boxes per record ~ clip(Poisson(4)+1, 1, 36)
(the data report's mean of 3.8), 30 candidates per query, 200 distinct queries.  Prints the rates of the two host stages
alone (decode; decode + assemble) and of the whole driver, so that the bound of the pipeline is visible.
`--host-only` stops before the scorer (no GPU needed).
"""
import base64
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import drivers, records, synth, tokenizer  # noqa: E402
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LDS, LXMERT, ZK, ModelConfig  # noqa: E402

LQ, R, BATCH = 32, 36, 256


def make_lines(n, vocab_words, n_labels, rng):
    queries = [" ".join(rng.choice(vocab_words, int(rng.integers(2, 7)))) for _ in range(200)]
    lines = []
    for i in range(n):
        nb = int(np.clip(rng.poisson(4) + 1, 1, R))
        h, w = int(rng.integers(200, 900)), int(rng.integers(200, 900))
        boxes = (np.sort(rng.random((nb, 2, 2)).astype(np.float32), axis=1) * np.array([h, w], np.float32)).reshape(nb, 4)
        feats = (np.abs(rng.standard_normal((nb, 2048))) * 0.5 * (rng.random((nb, 2048)) > 0.6)).astype(np.float32)
        labels = rng.integers(0, n_labels, nb).astype(np.int64)
        f = [str(i), str(h), str(w), str(nb), base64.b64encode(boxes.tobytes()).decode(),
             base64.b64encode(feats.tobytes()).decode(), base64.b64encode(labels.tobytes()).decode(),
             queries[(i // 30) % len(queries)], str(i // 30)]
        lines.append(("\t".join(f) + "\n").encode())
    return lines


def stage_times(sc, tok, label_map, lines):
    """Host time per call of the pipeline's stages inside one more score_tsv run (wrappers around the stage functions;
    waits on the GPU show up in the stage that blocks)."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import scorer as scorer_mod
    acc = {}

    def timed(name, fn):
        def wrapper(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                e = acc.setdefault(name, [0, 0.0])
                e[0] += 1
                e[1] += time.perf_counter() - t0
        return wrapper

    saved = (records.RecordDecoder.decode, records.FeedAssembler.assemble, records.normalize_boxes,
             scorer_mod.distinct_queries, scorer_mod.MatchScorer.forward_device, torch.cuda.Event.synchronize,
             drivers._ids_of)
    records.RecordDecoder.decode = timed("decode", saved[0])
    records.FeedAssembler.assemble = timed("assemble (incl. normalize_boxes)", saved[1])
    records.normalize_boxes = timed("normalize_boxes", saved[2])
    scorer_mod.distinct_queries = timed("distinct_queries", saved[3])
    scorer_mod.MatchScorer.forward_device = timed("forward_device (enqueue)", saved[4])
    torch.cuda.Event.synchronize = timed("wait: H2D of chunk i-2", saved[5])
    drivers._ids_of = timed("_ids_of", saved[6])
    try:
        t0 = time.perf_counter()
        drivers.score_tsv(sc, tok, label_map, lines)
        total = time.perf_counter() - t0
    finally:
        (records.RecordDecoder.decode, records.FeedAssembler.assemble, records.normalize_boxes,
         scorer_mod.distinct_queries, scorer_mod.MatchScorer.forward_device, torch.cuda.Event.synchronize,
         drivers._ids_of) = saved
    print(f"stage times of one score_tsv run ({total * 1e3:.1f} ms):")
    for name, (cnt, t) in acc.items():
        print(f"  {name:34s} {cnt:4d} calls  {t * 1e3:8.2f} ms total  {t / cnt * 1e3:7.3f} ms per call")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    host_only = "--host-only" in sys.argv
    n = int(args[0]) if args else 6144
    kind = args[1] if len(args) > 1 else ZK
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "tokenizer_kat.json"), encoding="utf-8"))
    vocab = {t: i for i, t in enumerate(dict.fromkeys(kat["vocab"]))}
    words = [t for t in vocab if t.isalpha() and t.isascii()]
    label_map = {i: " ".join(words[(3 * i + j) % len(words)] for j in range(1 + i % 3)) for i in range(33)}
    rng = np.random.default_rng(7)
    t0 = time.perf_counter()
    lines = make_lines(n, words, len(label_map), rng)
    mb = sum(map(len, lines)) / 1e6
    print(f"{n} lines, {mb:.0f} MB of TSV ({mb / n * 1e3:.0f} KB per line) generated in {time.perf_counter() - t0:.1f} s; "
          f"{os.cpu_count()} host threads", flush=True)
    depth = dict(n_layers=12) if kind != LXMERT else dict(n_layers=9, n_r_layers=5, n_x_layers=5)
    cfg = ModelConfig(kind, lq=LQ, nbox=R, vocab=max(vocab.values()) + 1, **depth)
    tok = tokenizer.FullTokenizer(vocab=vocab, max_input_chars_per_word=100 if kind == LXMERT else 200)

    # host stages alone
    dec = records.RecordDecoder(BATCH, max_boxes=R, feat_dim=cfg.feat_dim, pin=not host_only)
    asm = records.FeedAssembler(cfg, tok, label_map)
    for stage in ("decode", "decode + assemble (host part)"):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            for lo in range(0, n, BATCH):
                b = dec.decode(lines[lo:lo + BATCH])
                if stage != "decode" and kind == LDS:
                    asm.assemble(b)                     # (zk / lxmert normalise the boxes on the GPU: timed below)
            best = min(best, time.perf_counter() - t0)
        print(f"{stage:32s}: {n / best:9.0f} lines/s", flush=True)
        if kind != LDS:
            break
    if host_only:
        return

    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.scorer import MatchScorer
    sc = MatchScorer(cfg, synth.make_weights(cfg, seed=3), device=0, max_batch=BATCH)
    try:
        drivers.score_tsv(sc, tok, label_map, lines[:4 * BATCH])          # lazy init, graph capture, caches
        best, res = 1e9, None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = drivers.score_tsv(sc, tok, label_map, lines)
            best = min(best, time.perf_counter() - t0)
        assert np.isfinite(res["score"]).all() and len(res["score"]) == n
        if "--stages" in sys.argv:
            stage_times(sc, tok, label_map, lines)
        # the same pairs with the feeds already assembled and pinned (MatchScorer.score): what the driver adds on top
        feeds = records.FeedAssembler(cfg, tok, label_map).assemble(records.decode_lines(lines[:8 * BATCH], max_boxes=R))
        host = sc.to_feeds({k: v.cpu() for k, v in feeds.items()})
        sc.score(host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sc.score(host)
        t_feed = (time.perf_counter() - t0) / (8 * BATCH)
        print(json.dumps({"what": "drivers.score_tsv: TSV lines in memory -> scores (decode, assemble, H2D, forward, D2H)",
                          "model": kind, "lines": n, "tsv_mb": round(mb, 1), "pairs_per_s": round(n / best, 1),
                          "wall_ms": round(best * 1e3, 2), "score_from_pinned_feeds_pairs_per_s": round(1 / t_feed, 1),
                          "host_threads": os.cpu_count(), "shapes": f"{LQ} x {R} x 2048-d, batch {BATCH}"}))
    finally:
        sc.close()


if __name__ == "__main__":
    main()
