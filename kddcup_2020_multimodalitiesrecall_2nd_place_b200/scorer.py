"""Host side of the scorer: owns one mmr_handle (packed weights + workspace on one GPU) and moves batches through it.

PyTorch is used for device memory, pinned host memory and streams only; every arithmetic step is a kernel of
libmmrecall.so reached through the C ABI (include/mmrecall.h).  No CPU fallback exists: constructing a
MatchScorer without the library or without an sm_100 device raises.

Reference call sites replaced (the per-batch body of the three scoring drivers):
  imagebert_zk/evaluate_normal.py:222-249            sess.run([probs, total_loss], feed_dict)
  imagebert_lds/src/run_pretraining_predict_score.py:566-576
  lxmert/src/tasks/kdd_model.py:66-113               KDD.predict inner loop
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .config import KIND_CODE, LDS, LXMERT, ZK, ModelConfig

_DT = {"fp16": _lib.DT_FP16, "float16": _lib.DT_FP16, "bf16": _lib.DT_BF16, "bfloat16": _lib.DT_BF16}

# feed name -> (mmr_inputs member, torch dtype, trailing shape as a function of cfg)
_FEEDS = {
    ZK: ("query_ids", "segment_ids", "label_ids", "feats", "boxes", "len_query", "num_boxes", "labels"),
    LDS: ("query_ids", "segment_ids", "label_ids", "feats"),
    LXMERT: ("query_ids", "label_ids", "feats", "boxes", "query_mask", "visn_mask"),
}


def feed_spec(cfg: ModelConfig) -> Dict[str, tuple]:
    """name -> (torch dtype, per-pair shape) of every device input of `cfg.kind` (reference feeds:
    evaluate_normal.py:141-152, run_pretraining_predict_score.py:526-541, kdd_model.py:74-95)."""
    Lq, R, T = cfg.lq, cfg.nbox, cfg.label_len
    spec = {
        "query_ids": (torch.int32, (Lq,)),
        "label_ids": (torch.int32, (R, T)),
        "feats": (torch.float32, (R, cfg.feat_dim)),
    }
    if cfg.kind == ZK:
        spec.update(segment_ids=(torch.int32, (Lq + R,)), boxes=(torch.float32, (R, 5)), len_query=(torch.int32, ()),
                    num_boxes=(torch.int32, ()), labels=(torch.int32, ()))
    elif cfg.kind == LDS:
        spec.update(segment_ids=(torch.int32, (Lq,)))
    else:
        spec.update(boxes=(torch.float32, (R, 4)), query_mask=(torch.int32, (Lq,)), visn_mask=(torch.int32, (R,)))
    return {k: spec[k] for k in _FEEDS[cfg.kind]}


def distinct_queries(query_ids, query_mask):
    """LXMERT's first `n_layers` blocks see the query only (lxrt/modeling.py:577-578), and a candidate set scores ~30
    products per query: (lang_unique [U] int32 ascending representative pair indices, lang_slot [B] int32 position of
    every pair's representative in lang_unique) of a batch, for mmr_inputs.lang_unique / lang_slot.  Host arithmetic on
    [B, lq] integers."""
    q = query_ids.numpy() if torch.is_tensor(query_ids) else np.asarray(query_ids)
    m = query_mask.numpy() if torch.is_tensor(query_mask) else np.asarray(query_mask)
    key = np.ascontiguousarray(np.concatenate([q.astype(np.int32), m.astype(np.int32)], axis=1))
    rows = key.view(np.dtype((np.void, key.shape[1] * 4))).ravel()      # one opaque item per row: a 1-D unique
    _, first, inverse = np.unique(rows, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")              # representatives in ascending pair order
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return (torch.from_numpy(first[order].astype(np.int32)), torch.from_numpy(rank[inverse.reshape(-1)].astype(np.int32)))


class MatchScorer:
    """One model (zk / lds / lxmert) resident on one B200."""

    def __init__(self, cfg: ModelConfig, weights: Dict[str, np.ndarray], device: int = 0, dtype: str = "fp16",
                 max_batch: int = 256, precision: str = "fast"):
        """precision: "fast" = every MMA operand rounded once to `dtype`; "strict" = two-term split operands on every
        GEMM, precise activations, fp32 attention (~1e-5 of the fp32 reference path at about a third of the
        throughput; include/mmrecall.h, mmr_precision)."""
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        self.max_batch = int(max_batch)
        self.dtype = dtype
        if precision not in ("fast", "strict"):
            raise ValueError(f"precision must be 'fast' or 'strict', got {precision!r}")
        self.precision = precision
        _lib.check(self.lib.mmr_device_check(device))
        c = _lib.MmrConfig(
            model_kind=KIND_CODE[cfg.kind], dtype=_DT[dtype], hidden=cfg.hidden, heads=cfg.heads,
            intermediate=cfg.intermediate, vocab=cfg.vocab, max_pos=cfg.max_pos, type_vocab=cfg.type_vocab,
            feat_dim=cfg.feat_dim, label_len=cfg.label_len, n_layers=cfg.n_layers, n_r_layers=cfg.n_r_layers,
            n_x_layers=cfg.n_x_layers, lq=cfg.lq, nbox=cfg.nbox, max_batch=self.max_batch,
            precision=_lib.PRECISION_STRICT if precision == "strict" else _lib.PRECISION_FAST)
        keep = []
        arr = (_lib.MmrTensor * len(weights))()
        for i, (name, w) in enumerate(weights.items()):
            a = np.ascontiguousarray(w.detach().cpu().numpy() if torch.is_tensor(w) else w, dtype=np.float32)
            keep.append(a)
            arr[i].name = name.encode()
            arr[i].data = a.ctypes.data
            arr[i].ndim = min(a.ndim, 4)
            dims = list(a.shape) if a.ndim <= 4 else [int(np.prod(a.shape[:-3]))] + list(a.shape[-3:])
            for j, d in enumerate(dims):
                arr[i].dims[j] = d
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mmr_create(C.byref(c), arr, len(weights), device, C.byref(h)))
        self._h = h
        self.spec = feed_spec(cfg)
        self._slots = None
        # CUDA graphs of the forward, keyed on (batch, every pointer, tuning generation): the second forward on the
        # same buffers is captured, later ones are replayed (the 69 launches of a forward then cost one graph launch;
        # +2 % on the 12-layer forward).  mmr_forward neither allocates nor synchronises and keeps its state on the
        # device, which is what makes the replay exact (tests/test_gpu_parity.py).  MMR_CUDA_GRAPHS=0 turns it off.
        self.use_graphs = os.environ.get("MMR_CUDA_GRAPHS", "1") != "0"
        # LXMERT: evaluate the query-only language blocks once per distinct query of a batch (mmr_inputs.lang_unique);
        # to_feeds / score_stream derive the grouping from the host copies of query_ids / query_mask
        self.dedup_queries = cfg.kind == LXMERT
        self._graphs, self._seen = {}, set()
        self._capture_stream = None
        self._taps_on = self._profiling_on = False

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_graphs", None):
            self._graphs.clear()               # captured forwards point into the workspace that goes away below
            self._seen.clear()
        if getattr(self, "_h", None):
            self.lib.mmr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ device-resident path
    def _check_feed(self, name, t, B):
        dt, shape = self.spec[name]
        if t.dtype != dt or tuple(t.shape) != (B, *shape) or not t.is_contiguous():
            raise ValueError(f"feed '{name}': need contiguous {dt} {(B, *shape)}, got {t.dtype} {tuple(t.shape)}")

    def forward_device(self, feeds: Dict[str, torch.Tensor], probs_out: Optional[torch.Tensor] = None,
                       pooled_out: Optional[torch.Tensor] = None,
                       logits_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Scores B <= max_batch pairs whose feeds already live on this GPU; asynchronous on the current stream.
        Returns probs [B,2] fp32 (the reference score is probs[:, 1])."""
        B = feeds["query_ids"].shape[0]
        inp = _lib.MmrInputs()
        fused = feeds.get("region_sum") if self.cfg.kind == ZK else None
        for name in self.spec:
            if fused is not None and name in ("feats", "boxes", "label_ids"):
                continue
            t = feeds[name]
            self._check_feed(name, t, B)
            setattr(inp, name, t.data_ptr())
        if fused is not None:
            if fused.dtype != torch.float32 or tuple(fused.shape) != (B, self.cfg.nbox, self.cfg.hidden) \
                    or not fused.is_contiguous():
                raise ValueError("feed 'region_sum': need contiguous float32 [B, nbox, hidden]")
            inp.region_sum = fused.data_ptr()
        n_unique = 0
        if self.cfg.kind == LXMERT and feeds.get("lang_unique") is not None and feeds.get("lang_slot") is not None:
            lu, ls = feeds["lang_unique"], feeds["lang_slot"]
            if lu.dtype != torch.int32 or ls.dtype != torch.int32 or tuple(ls.shape) != (B,) or lu.dim() != 1 \
                    or not (0 < lu.shape[0] <= B) or not lu.is_cuda or not ls.is_cuda:
                raise ValueError("feeds 'lang_unique' [U] / 'lang_slot' [B]: need int32 device tensors, 0 < U <= B")
            n_unique = int(lu.shape[0])
            inp.lang_unique, inp.lang_slot, inp.n_lang_unique = lu.data_ptr(), ls.data_ptr(), n_unique
        caller_owns_output = probs_out is not None
        if probs_out is None:
            probs_out = torch.empty((B, 2), dtype=torch.float32, device=self.device)

        def launch():
            _lib.check(self.lib.mmr_forward(self._h, C.byref(inp), B, probs_out.data_ptr(),
                                            0 if logits_out is None else logits_out.data_ptr(),
                                            0 if pooled_out is None else pooled_out.data_ptr(),
                                            torch.cuda.current_stream(self.device).cuda_stream))

        if not (self.use_graphs and caller_owns_output and not self._taps_on and not self._profiling_on) \
                or torch.cuda.is_current_stream_capturing():
            launch()
            return probs_out
        key = (B, n_unique, int(self.lib.mmr_tuning_generation()), probs_out.data_ptr(),
               0 if logits_out is None else logits_out.data_ptr(), 0 if pooled_out is None else pooled_out.data_ptr(),
               tuple(int(getattr(inp, f[0]) or 0) for f in inp._fields_))
        g = self._graphs.get(key)
        if g is None:
            if key not in self._seen or len(self._graphs) >= 64:
                if len(self._seen) > 4096:     # a caller that never reuses buffers: do not grow without bound
                    self._seen.clear()
                self._seen.add(key)            # first forward on these buffers: eager (also warms every lazy init)
                launch()
                return probs_out
            if self._capture_stream is None:
                self._capture_stream = torch.cuda.Stream(self.device)
            cur = torch.cuda.current_stream(self.device)
            self._capture_stream.wait_stream(cur)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(self._capture_stream):
                with torch.cuda.graph(g, stream=self._capture_stream):
                    launch()
            cur.wait_stream(self._capture_stream)
            self._graphs[key] = g
        g.replay()
        return probs_out

    def launches_per_forward(self) -> int:
        return int(self.lib.mmr_launches_per_forward(self._h))

    def set_profiling(self, on: bool):
        """Per-launch CUDA-event timing inside forward_device (bench.py's roofline line)."""
        self._profiling_on = bool(on)
        _lib.check(self.lib.mmr_set_profiling(self._h, int(on)))

    def profile(self):
        """[(kind, ms, flops)] of the last forward; kind 0 GEMM, 1 attention, 2 LayerNorm, 3 embed/head rows."""
        cap = 512
        kinds, ms, fl = (C.c_int32 * cap)(), (C.c_float * cap)(), (C.c_double * cap)()
        n = self.lib.mmr_get_profile(self._h, cap, kinds, ms, fl)
        if n < 0:
            raise _lib.MmrError("mmr_get_profile failed")
        return [(int(kinds[i]), float(ms[i]), float(fl[i])) for i in range(n)]

    def set_debug_taps(self, level):
        """0 off, 1 keep the embedding output, 2 also keep every encoder layer's output (single-stream models)."""
        self._taps_on = int(level) != 0
        _lib.check(self.lib.mmr_set_debug_taps(self._h, int(level)))

    def activation(self, which: int, batch: int) -> torch.Tensor:
        """Parity tap after the last forward: 0 = embedding output, 1 = final encoder output, 2 + i = encoder layer i;
        [rows, hidden]."""
        cfg = self.cfg
        rows = batch * (cfg.lq + (2 if cfg.kind == LDS else 1) * cfg.nbox)
        out = torch.empty((rows, cfg.hidden), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.mmr_get_activation(self._h, which, out.data_ptr(), out.numel(),
                                               torch.cuda.current_stream(self.device).cuda_stream))
        return out

    # ------------------------------------------------------------------ host-buffer path (what a driver calls)
    def to_feeds(self, arrays: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
        """Host arrays (any int width / float32) -> CPU tensors in the exact feed dtypes, pinned."""
        out = {}
        limits = {"query_ids": self.cfg.vocab, "label_ids": self.cfg.vocab, "segment_ids": self.cfg.type_vocab}
        for name, (dt, _) in self.spec.items():
            a = arrays[name]
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            if name in limits and t.numel() > 0:
                # ids index the embedding tables on the device: an id outside the bound checkpoint's vocabulary (a
                # vocab.txt that does not belong to it) would be an out-of-bounds gather, so it is refused here
                lo, hi = int(t.min()), int(t.max())
                if lo < 0 or hi >= limits[name]:
                    raise ValueError(f"feed '{name}': ids must lie in [0, {limits[name]}), got [{lo}, {hi}]")
            t = t.to(dt).contiguous()
            out[name] = t if t.is_pinned() else t.pin_memory()
        if self.cfg.kind == LXMERT and self.dedup_queries and out["query_ids"].shape[0] > 1:
            out["lang_unique"], out["lang_slot"] = distinct_queries(out["query_ids"], out["query_mask"])
        return out

    def _make_slots(self):
        if self._slots is None:
            Bm = self.max_batch
            self._slots = []
            for _ in range(2):
                dev = {n: torch.empty((Bm, *shape), dtype=dt, device=self.device) for n, (dt, shape) in self.spec.items()}
                host = {}
                if self.cfg.kind == LXMERT:
                    for n in ("lang_unique", "lang_slot"):
                        dev[n] = torch.empty((Bm,), dtype=torch.int32, device=self.device)
                        host[n] = torch.empty((Bm,), dtype=torch.int32).pin_memory()
                self._slots.append({"dev": dev, "host": host, "free": torch.cuda.Event(), "ready": torch.cuda.Event(),
                                    "probs": torch.empty((Bm, 2), dtype=torch.float32, device=self.device)})
            self._copy_stream = torch.cuda.Stream(self.device)
        return self._slots

    @property
    def copy_stream(self) -> "torch.cuda.Stream":
        """The stream score_stream issues its input copies on.  A `fetch` callback that prepares part of a chunk ON THE
        DEVICE (records.normalize_boxes) must do so on this stream, so that the copy into the slot is ordered after it
        without stalling the copy pipeline behind the compute stream."""
        self._make_slots()
        return self._copy_stream

    def score(self, feeds_host: Dict[str, torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Scores N pairs held in (ideally pinned) host tensors: N is cut into max_batch chunks whose H2D copies run on
        a copy stream, double-buffered against the kernels of the previous chunk; probabilities come back to pinned
        host memory.  Returns probs [N,2] (CPU, pinned); synchronises once at the end."""
        N = feeds_host["query_ids"].shape[0]
        return self.score_stream(N, lambda lo, hi: {n: feeds_host[n][lo:hi] for n in self.spec}, out)

    def score_stream(self, n_pairs: int, fetch, out: Optional[torch.Tensor] = None, on_device: bool = False) -> torch.Tensor:
        """The same pipeline over a pair list that need not be resident as one array: `fetch(lo, hi)` returns the feeds
        of pairs [lo, hi) (hi - lo <= max_batch; a decoder's batch arrays, a window of a memory-mapped file, ...; host or
        device tensors).  `fetch` may reuse its buffers every SECOND call: before call k, the copies that read what call
        k - 2 returned have completed.  on_device=True returns the [N, 2] probabilities as a DEVICE tensor, asynchronously
        on the current stream (no host copy, no synchronisation): what the sharded path gathers from."""
        N = int(n_pairs)
        if on_device:
            if N == 0:
                return torch.empty((0, 2), dtype=torch.float32, device=self.device)
        else:
            if out is None:
                out = torch.empty((N, 2), dtype=torch.float32).pin_memory()
            if N == 0:
                return out
        slots = self._make_slots()
        compute = torch.cuda.current_stream(self.device)
        copy = self._copy_stream
        copy.wait_stream(compute)
        dev_probs = torch.empty((N, 2), dtype=torch.float32, device=self.device)
        Bm = self.max_batch
        for i, lo in enumerate(range(0, N, Bm)):
            hi = min(N, lo + Bm)
            s = slots[i % 2]
            if i >= 2:
                s["ready"].synchronize()                  # H2D of chunk i-2 done: its source buffers may be rewritten
            chunk = fetch(lo, hi)
            group = None
            if self.cfg.kind == LXMERT and self.dedup_queries and hi - lo > 1 and not chunk["query_ids"].is_cuda \
                    and not chunk["query_mask"].is_cuda:
                # (the slot's pinned staging tensors: reusable, the copies that read them two chunks ago have completed)
                uniq, slot_of = distinct_queries(chunk["query_ids"], chunk["query_mask"])
                s["host"]["lang_unique"][: uniq.shape[0]].copy_(uniq)
                s["host"]["lang_slot"][: hi - lo].copy_(slot_of)
                group = (s["host"]["lang_unique"][: uniq.shape[0]], s["host"]["lang_slot"][: hi - lo])
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(s["free"])            # kernels of chunk i-2 have consumed this slot
                for name in self.spec:
                    src = chunk[name]
                    if not src.is_cuda and not src.is_pinned():
                        # a copy from pageable memory blocks the host until the copy stream gets to it (i.e. until the
                        # kernels two chunks back are done): go through the slot's pinned staging instead
                        stage = s["host"].get(name)
                        if stage is None:
                            dt, shape = self.spec[name]
                            stage = s["host"][name] = torch.empty((Bm, *shape), dtype=dt).pin_memory()
                        # (numpy: a torch copy_ of this size wakes the intra-op thread pool, whose workers then spin
                        # on every core for a while -- under the feet of a driver's decoder threads)
                        np.copyto(stage.numpy()[: hi - lo], src.detach().numpy(), casting="unsafe")
                        src = stage[: hi - lo]
                    s["dev"][name][: hi - lo].copy_(src, non_blocking=True)
                if group is not None:
                    s["dev"]["lang_unique"][: group[0].shape[0]].copy_(group[0], non_blocking=True)
                    s["dev"]["lang_slot"][: hi - lo].copy_(group[1], non_blocking=True)
                s["ready"].record(copy)
            compute.wait_event(s["ready"])
            # into the slot's own probs buffer: the forward then sees the same pointers every other chunk (graph replay)
            dev_feeds = {n: s["dev"][n][: hi - lo] for n in self.spec}
            if group is not None:
                dev_feeds["lang_unique"] = s["dev"]["lang_unique"][: group[0].shape[0]]
                dev_feeds["lang_slot"] = s["dev"]["lang_slot"][: hi - lo]
            self.forward_device(dev_feeds, probs_out=s["probs"][: hi - lo])
            dev_probs[lo:hi].copy_(s["probs"][: hi - lo], non_blocking=True)
            s["free"].record(compute)
        if on_device:
            return dev_probs
        out.copy_(dev_probs, non_blocking=True)
        compute.synchronize()
        return out


def sharded_score(scorer: MatchScorer, feeds_host: Dict[str, torch.Tensor], rank: int, world: int,
                  gather: bool = True) -> torch.Tensor:
    """Embarrassingly parallel scoring of N pairs over `world` ranks (one process per GPU): rank r scores the
    contiguous range [r*ceil(N/W), (r+1)*ceil(N/W)) and the fp32 scores are concatenated with ONE all-gather
    (NCCL over NVLink when the process group is nccl; gloo in the CPU tests of the host logic).  Returns the full
    [N] score vector on every rank (CPU tensor)."""
    N = feeds_host["query_ids"].shape[0]
    return sharded_score_stream(scorer, N, lambda lo, hi: {k: feeds_host[k][lo:hi] for k in scorer.spec}, rank, world,
                                gather)


def sharded_score_stream(scorer: MatchScorer, n_pairs: int, fetch, rank: int, world: int,
                         gather: bool = True) -> torch.Tensor:
    """sharded_score over a pair list given by `fetch(lo, hi)` in GLOBAL pair indices: every rank touches only the
    host feeds of its own range (at cfg4 a rank stages 1.1 GB instead of the whole 8.8 GB candidate set).  Under NCCL
    the shard's scores go from the scorer's device buffer straight into the all-gather -- no host round trip."""
    import torch.distributed as dist
    lo, hi, per = shard_range(n_pairs, rank, world)
    local = lambda a, b: fetch(lo + a, lo + b)
    if gather and world > 1 and dist.is_initialized() and dist.get_backend() == "nccl" and hasattr(scorer, "device"):
        probs = scorer.score_stream(hi - lo, local, on_device=True)
        mine = torch.zeros(per, dtype=torch.float32, device=scorer.device)
        mine[: hi - lo] = probs[:, 1]
        out = torch.empty(world * per, dtype=torch.float32, device=scorer.device)
        dist.all_gather_into_tensor(out, mine)
        return out[:n_pairs].cpu()
    probs = scorer.score_stream(hi - lo, local) if hi > lo else torch.empty((0, 2))
    mine = torch.zeros(per, dtype=torch.float32)
    mine[: hi - lo] = probs[:, 1]
    if not gather or world == 1:
        return mine[: hi - lo] if world == 1 else mine
    return allgather_scores(mine, n_pairs, world, scorer.device)


def shard_range(n: int, rank: int, world: int):
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def allgather_scores(mine: torch.Tensor, n: int, world: int, device=None) -> torch.Tensor:
    """The only collective of the design: all-gather of per-rank fp32 score shards (padded to equal length)."""
    import torch.distributed as dist
    backend = dist.get_backend()
    buf = mine.to(device) if backend == "nccl" else mine
    out = torch.empty(world * buf.numel(), dtype=torch.float32, device=buf.device)
    dist.all_gather_into_tensor(out, buf)
    return out[:n].cpu() if backend == "nccl" else out[:n]
