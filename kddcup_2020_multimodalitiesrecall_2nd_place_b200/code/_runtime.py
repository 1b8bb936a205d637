"""Plumbing shared by the reference-named modules: weight binding, scorer cache, chunked execution."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib
from ..config import LDS, LXMERT, ZK, ModelConfig
from ..scorer import MatchScorer

_BOUND: Dict[str, dict] = {}
MAX_BATCH = 256


def bind(kind: str, weights: Dict[str, np.ndarray], device: int = 0, dtype: str = "fp16", precision: str = "fast",
         **layers) -> None:
    """Registers the checkpoint of one model kind (names as in the reference checkpoints).
    `layers` overrides the depth (n_layers / n_r_layers / n_x_layers), e.g. for the 2-layer plumbing config;
    precision="strict" selects the two-term split-operand arithmetic (scorer.MatchScorer)."""
    release(kind)
    _BOUND[kind] = {"weights": weights, "device": device, "dtype": dtype, "precision": precision, "layers": layers,
                    "scorers": {}}


def release(kind: Optional[str] = None) -> None:
    for k in ([kind] if kind else list(_BOUND)):
        b = _BOUND.pop(k, None)
        if b:
            for sc in b["scorers"].values():
                sc.close()


def bound(kind: str) -> dict:
    if kind not in _BOUND:
        raise RuntimeError(f"no weights bound for '{kind}': call bind(weights) first (it replaces saver.restore / "
                           f"load_state_dict of the reference driver)")
    return _BOUND[kind]


def _depth(kind: str, weights, layers: dict) -> dict:
    """Layer counts from the checkpoint's own variable names unless given."""
    if kind == LXMERT:
        def count(tag):
            n = 0
            while f"lxrt_encoder.model.bert.encoder.{tag}.{n}.attention.self.query.weight" in weights or \
                    f"lxrt_encoder.model.bert.encoder.{tag}.{n}.visual_attention.att.query.weight" in weights:
                n += 1
            return n
        d = {"n_layers": count("layer"), "n_r_layers": count("r_layers"), "n_x_layers": count("x_layers")}
    else:
        n = 0
        while f"bert/encoder/layer_{n}/attention/self/query/kernel" in weights:
            n += 1
        d = {"n_layers": n}
    d.update(layers)
    return d


def scorer_for(kind: str, lq: int, nbox: int, batch: int) -> MatchScorer:
    b = bound(kind)
    mb = min(MAX_BATCH, max(batch, 1))
    for (klq, knb, kmb), sc in b["scorers"].items():
        if klq == lq and knb == nbox and kmb >= mb:
            return sc
    w = b["weights"]
    vocab = w["bert/embeddings/word_embeddings" if kind != LXMERT else
              "lxrt_encoder.model.bert.embeddings.word_embeddings.weight"].shape[0]
    cfg = ModelConfig(kind, lq=lq, nbox=nbox, vocab=int(vocab), **_depth(kind, w, b["layers"]))
    sc = MatchScorer(cfg, w, device=b["device"], dtype=b["dtype"], max_batch=mb, precision=b.get("precision", "fast"))
    b["scorers"][(lq, nbox, mb)] = sc
    return sc


def as_tensor(a, dtype) -> torch.Tensor:
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(dtype).contiguous()


def run(sc: MatchScorer, feeds: Dict[str, torch.Tensor], pooled=False, logits=False, sequence=False, embedding=False,
        all_layers=False) -> Dict[str, torch.Tensor]:
    """Runs B pairs in chunks of the scorer's max_batch; feeds may live on the host or on the scorer's GPU.
    Returns device tensors: probs [B,2] and, on request, logits [B,2], pooled [B,H], sequence / embedding
    [B,S,H] (LXMERT: language rows then visual rows per chunk, see MatchScorer.activation), all_layers list."""
    B = feeds["query_ids"].shape[0]
    dev = sc.device
    H = sc.cfg.hidden
    out = {"probs": torch.empty((B, 2), dtype=torch.float32, device=dev)}
    if pooled:
        out["pooled"] = torch.empty((B, H), dtype=torch.float32, device=dev)
    if logits:
        out["logits"] = torch.empty((B, 2), dtype=torch.float32, device=dev)
    seq_chunks, emb_chunks, layer_chunks = [], [], []
    sc.set_debug_taps(2 if all_layers else (1 if (embedding or sequence) else 0))   # taps: every row of the last block
    for lo in range(0, B, sc.max_batch):
        hi = min(B, lo + sc.max_batch)
        chunk = {k: v[lo:hi].to(dev, non_blocking=True) for k, v in feeds.items()}
        sc.forward_device(chunk, probs_out=out["probs"][lo:hi],
                          pooled_out=out["pooled"][lo:hi] if pooled else None,
                          logits_out=out["logits"][lo:hi] if logits else None)
        if sequence:
            seq_chunks.append(sc.activation(1, hi - lo))
        if embedding:
            emb_chunks.append(sc.activation(0, hi - lo))
        if all_layers:
            layer_chunks.append([sc.activation(2 + i, hi - lo) for i in range(sc.cfg.n_layers)])
    sc.set_debug_taps(0)
    if sequence:
        out["sequence"] = seq_chunks
    if embedding:
        out["embedding"] = emb_chunks
    if all_layers:
        out["all_layers"] = layer_chunks
    return out


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor):
    """(mean loss, per-example loss, log_probs) of a 2-way head; O(B) device arithmetic on the kernel's logits."""
    log_probs = torch.log_softmax(logits, dim=-1)
    per_example = -log_probs.gather(1, labels.to(logits.device).long().view(-1, 1)).squeeze(1)
    return per_example.mean(), per_example, log_probs


def am_softmax_head(pooled: torch.Tensor, am_kernel: np.ndarray, labels: torch.Tensor):
    """model_triple.amsoftmax_loss (model_triple.py:56-86) on device rows: returns (probs, logits)."""
    lib = _lib.load()
    k = np.asarray(am_kernel, np.float32)                                  # [768, 2]
    kn = k / np.sqrt(np.maximum((k * k).sum(0, keepdims=True), 1e-10))     # l2_normalize(kernel, 0, 1e-10)
    wn = torch.from_numpy(np.ascontiguousarray(kn.T)).to(pooled.device)    # [2, 768]
    B = pooled.shape[0]
    probs = torch.empty((B, 2), dtype=torch.float32, device=pooled.device)
    logits = torch.empty((B, 2), dtype=torch.float32, device=pooled.device)
    lab = labels.to(pooled.device).to(torch.int32).contiguous()
    x = pooled.float().contiguous()
    _lib.check(lib.mmr_am_softmax_head(x.data_ptr(), wn.data_ptr(), lab.data_ptr(), B, probs.data_ptr(),
                                       logits.data_ptr(), torch.cuda.current_stream(pooled.device).cuda_stream))
    return probs, logits


def linear_head(x: torch.Tensor, W: np.ndarray, bias: np.ndarray, ln_gamma=None, ln_beta=None):
    """softmax(LN?(x) . W^T + b) for W [2, width]: returns (probs, logits)."""
    lib = _lib.load()
    dev = x.device
    x = x.float().contiguous()
    B, width = x.shape
    Wd = torch.from_numpy(np.ascontiguousarray(W, dtype=np.float32)).to(dev)
    bd = torch.from_numpy(np.ascontiguousarray(bias, dtype=np.float32)).to(dev)
    g = None if ln_gamma is None else torch.from_numpy(np.ascontiguousarray(ln_gamma, dtype=np.float32)).to(dev)
    be = None if ln_beta is None else torch.from_numpy(np.ascontiguousarray(ln_beta, dtype=np.float32)).to(dev)
    probs = torch.empty((B, 2), dtype=torch.float32, device=dev)
    logits = torch.empty((B, 2), dtype=torch.float32, device=dev)
    _lib.check(lib.mmr_linear_head(x.data_ptr(), width, 0 if g is None else g.data_ptr(),
                                   0 if be is None else be.data_ptr(), Wd.data_ptr(), bd.data_ptr(), B,
                                   probs.data_ptr(), logits.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return probs, logits


def prefix_lengths(mask: torch.Tensor, name: str) -> torch.Tensor:
    """Lengths of a tf.sequence_mask-style mask [B, L]; anything that is not a prefix of ones is rejected (the
    kernels rebuild the key mask from lengths, evaluate_normal.py feeds lengths, model_triple.py:198-199)."""
    m = mask.to(torch.int64)
    n = m.sum(1)
    want = (torch.arange(m.shape[1])[None, :] < n[:, None]).to(torch.int64)
    if not torch.equal(m.cpu(), want.cpu()):
        raise ValueError(f"{name}: only prefix (sequence_mask) masks are supported")
    return n.to(torch.int32)
