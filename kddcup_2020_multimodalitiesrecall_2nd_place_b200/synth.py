"""Seeded synthetic weights and inputs (no checkpoints or competition data ship with the reference).

Weights are named exactly as in the reference checkpoints and follow the reference
initialisers: TF models truncated-normal(0.02) (pixelbert.py:427-429), xavier-normal am_kernel
(model_triple.py:62-63); torch LXMERT normal(0.02), LayerNorm 1/0 (modeling.py:715-726).  `trained_like=True`
perturbs them (matrices x3, LN gamma ~ U(0.5,1.5), beta ~ N(0,0.1), biases ~ N(0,0.02)) to stress the tolerance.

Inputs follow the loaders' layouts: load_data_v4.py:133-163, 204, 264-265, 380-389 (zk);
load_data_pred.py:94-121, 159, 211-241 (lds); lxmert/src/utils.py:23-59, tasks/kdd_data.py:39-125.
"""
from __future__ import annotations

import numpy as np

from .config import LDS, LXMERT, ZK, ModelConfig

SEED0 = 20200823


def _tn(rng, shape, std=0.02):
    """truncated normal at 2 sigma (tf.truncated_normal_initializer)."""
    x = rng.standard_normal(size=shape).astype(np.float32)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(size=int(bad.sum())).astype(np.float32)
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


class _Init:
    def __init__(self, rng, trained_like, truncated):
        self.rng, self.tl, self.trunc = rng, trained_like, truncated

    def mat(self, *shape, std=0.02):
        w = _tn(self.rng, shape, std) if self.trunc else (self.rng.standard_normal(size=shape) * std).astype(np.float32)
        return (w * 3.0).astype(np.float32) if self.tl else w

    def emb(self, *shape):
        return _tn(self.rng, shape) if self.trunc else (self.rng.standard_normal(size=shape) * 0.02).astype(np.float32)

    def bias(self, n):
        return (self.rng.standard_normal(n) * 0.02).astype(np.float32) if self.tl else np.zeros(n, np.float32)

    def gamma(self, n):
        return self.rng.uniform(0.5, 1.5, n).astype(np.float32) if self.tl else np.ones(n, np.float32)

    def beta(self, n):
        return (self.rng.standard_normal(n) * 0.1).astype(np.float32) if self.tl else np.zeros(n, np.float32)

    def xavier(self, fan_in, fan_out):
        std = np.sqrt(2.0 / (fan_in + fan_out))
        return (self.rng.standard_normal((fan_in, fan_out)) * std).astype(np.float32)


def _tf_bert_tree(w, ini, cfg: ModelConfig):
    H, I = cfg.hidden, cfg.intermediate
    w["bert/embeddings/word_embeddings"] = ini.emb(cfg.vocab, H)
    w["bert/embeddings/token_type_embeddings"] = ini.emb(cfg.type_vocab, H)
    w["bert/embeddings/position_embeddings"] = ini.emb(cfg.max_pos, H)
    w["bert/embeddings/LayerNorm/gamma"] = ini.gamma(H)
    w["bert/embeddings/LayerNorm/beta"] = ini.beta(H)
    for i in range(cfg.n_layers):
        p = f"bert/encoder/layer_{i}/"
        for n in ("query", "key", "value"):
            w[p + f"attention/self/{n}/kernel"] = ini.mat(H, H)
            w[p + f"attention/self/{n}/bias"] = ini.bias(H)
        w[p + "attention/output/dense/kernel"] = ini.mat(H, H)
        w[p + "attention/output/dense/bias"] = ini.bias(H)
        w[p + "attention/output/LayerNorm/gamma"] = ini.gamma(H)
        w[p + "attention/output/LayerNorm/beta"] = ini.beta(H)
        w[p + "intermediate/dense/kernel"] = ini.mat(H, I)
        w[p + "intermediate/dense/bias"] = ini.bias(I)
        w[p + "output/dense/kernel"] = ini.mat(I, H)
        w[p + "output/dense/bias"] = ini.bias(H)
        w[p + "output/LayerNorm/gamma"] = ini.gamma(H)
        w[p + "output/LayerNorm/beta"] = ini.beta(H)
    w["bert/pooler/dense/kernel"] = ini.mat(H, H)
    w["bert/pooler/dense/bias"] = ini.bias(H)


def make_weights(cfg: ModelConfig, seed: int = SEED0, trained_like: bool = False) -> dict:
    """name -> fp32 ndarray, reference checkpoint names and layouts (TF kernels [in,out]; torch [out,in])."""
    rng = np.random.default_rng(seed)
    H, I, F = cfg.hidden, cfg.intermediate, cfg.feat_dim
    w = {}
    if cfg.kind == ZK:
        ini = _Init(rng, trained_like, truncated=True)
        _tf_bert_tree(w, ini, cfg)
        # slim xavier init for the kdd_* layers is replaced by the same N(0,0.02) family (documented synthetic choice)
        w["kdd_conv1/weights"] = ini.mat(1, cfg.label_len, H, H)
        w["kdd_conv1/biases"] = ini.bias(H)
        w["kdd_dense1/weights"] = ini.mat(5, H, std=0.2)
        w["kdd_dense1/biases"] = ini.bias(H)
        w["kdd_conv2/weights"] = ini.mat(1, 1, F, H)
        w["kdd_conv2/biases"] = ini.bias(H)
        w["kdd_featureemb/fully_connected/weights"] = ini.mat(H, H)
        w["kdd_featureemb/fully_connected/biases"] = ini.bias(H)
        w["cls/seq_relationship/am_kernel"] = ini.xavier(H, 2)
    elif cfg.kind == LDS:
        ini = _Init(rng, trained_like, truncated=True)
        _tf_bert_tree(w, ini, cfg)
        w["featureemb/fully_connected/weights"] = ini.mat(F, H)
        w["featureemb/fully_connected/biases"] = ini.bias(H)
        w["bert/embeddings/word_embeddings_labelembedding"] = ini.emb(cfg.label_len, 1)
        w["cls/seq_relationship/output_weights"] = ini.mat(2, H)
        w["cls/seq_relationship/output_bias"] = ini.bias(2)
    elif cfg.kind == LXMERT:
        ini = _Init(rng, trained_like, truncated=False)
        b = "lxrt_encoder.model.bert."
        w[b + "embeddings.word_embeddings.weight"] = ini.emb(cfg.vocab, H)
        w[b + "embeddings.position_embeddings.weight"] = ini.emb(cfg.max_pos, H)
        w[b + "embeddings.token_type_embeddings.weight"] = ini.emb(cfg.type_vocab, H)
        w[b + "embeddings.LayerNorm.weight"] = ini.gamma(H)
        w[b + "embeddings.LayerNorm.bias"] = ini.beta(H)
        v = b + "encoder.visn_fc."
        w[v + "visn_fc.weight"] = ini.mat(H, F)
        w[v + "visn_fc.bias"] = ini.bias(H)
        w[v + "visn_layer_norm.weight"] = ini.gamma(H)
        w[v + "visn_layer_norm.bias"] = ini.beta(H)
        w[v + "box_fc.weight"] = ini.mat(H, 4, std=0.2)
        w[v + "box_fc.bias"] = ini.bias(H)
        w[v + "box_layer_norm.weight"] = ini.gamma(H)
        w[v + "box_layer_norm.bias"] = ini.beta(H)
        w[v + "label_conv.weight"] = (rng.standard_normal((1, cfg.label_len, 1, 1)) * 0.3).astype(np.float32)
        w[v + "label_conv.bias"] = (rng.standard_normal(1) * 0.05).astype(np.float32)
        w[v + "label_fc.weight"] = ini.mat(H, H)
        w[v + "label_fc.bias"] = ini.bias(H)
        w[v + "label_layer_norm.weight"] = ini.gamma(H)
        w[v + "label_layer_norm.bias"] = ini.beta(H)

        def att(p):
            for n in ("query", "key", "value"):
                w[p + f"{n}.weight"] = ini.mat(H, H)
                w[p + f"{n}.bias"] = ini.bias(H)

        def att_out(p):
            w[p + "dense.weight"] = ini.mat(H, H)
            w[p + "dense.bias"] = ini.bias(H)
            w[p + "LayerNorm.weight"] = ini.gamma(H)
            w[p + "LayerNorm.bias"] = ini.beta(H)

        def ffn(pi, po):
            w[pi + "dense.weight"] = ini.mat(I, H)
            w[pi + "dense.bias"] = ini.bias(I)
            w[po + "dense.weight"] = ini.mat(H, I)
            w[po + "dense.bias"] = ini.bias(H)
            w[po + "LayerNorm.weight"] = ini.gamma(H)
            w[po + "LayerNorm.bias"] = ini.beta(H)

        for group, n in (("layer", cfg.n_layers), ("r_layers", cfg.n_r_layers)):
            for i in range(n):
                p = b + f"encoder.{group}.{i}."
                att(p + "attention.self.")
                att_out(p + "attention.output.")
                ffn(p + "intermediate.", p + "output.")
        for i in range(cfg.n_x_layers):
            p = b + f"encoder.x_layers.{i}."
            att(p + "visual_attention.att.")
            att_out(p + "visual_attention.output.")
            for s in ("lang", "visn"):
                att(p + f"{s}_self_att.self.")
                att_out(p + f"{s}_self_att.output.")
                ffn(p + f"{s}_inter.", p + f"{s}_output.")
        w[b + "pooler.dense.weight"] = ini.mat(H, H)
        w[b + "pooler.dense.bias"] = ini.bias(H)
        w["logit_fc.0.weight"] = ini.mat(2 * H, H)
        w["logit_fc.0.bias"] = ini.bias(2 * H)
        w["logit_fc.2.weight"] = ini.gamma(2 * H)
        w["logit_fc.2.bias"] = ini.beta(2 * H)
        w["logit_fc.3.weight"] = ini.mat(2, 2 * H)
        w["logit_fc.3.bias"] = ini.bias(2)
    else:
        raise ValueError(cfg.kind)
    return w


def make_inputs(cfg: ModelConfig, batch: int, seed: int = SEED0, n_queries: int | None = None) -> dict:
    """Numpy feeds for `batch` pairs.  `n_queries`: pairs share query ids in contiguous groups (testB shape:
    each query is scored against ~30 candidates); default = every pair has its own query."""
    rng = np.random.default_rng(seed + 7919)
    B, Lq, R, T = batch, cfg.lq, cfg.nbox, cfg.label_len
    V = cfg.vocab
    lo_id = min(1000, V // 2)

    nq = B if n_queries is None else n_queries
    q_ids = np.zeros((nq, Lq), np.int32)
    q_len = np.clip(rng.poisson(4, nq) + 3, 3, Lq).astype(np.int32)
    for i in range(nq):
        n = int(q_len[i])
        q_ids[i, 0] = min(101, V - 1)
        q_ids[i, 1:n - 1] = rng.integers(lo_id, V, n - 2)
        q_ids[i, n - 1] = min(102, V - 1)
    owner = (np.arange(B) * nq) // B
    query_ids = q_ids[owner]
    len_query = q_len[owner]

    num_boxes = np.clip(rng.poisson(4, B) + 1, 1, R).astype(np.int32)
    feats = (np.abs(rng.standard_normal((B, R, cfg.feat_dim))) * 0.5).astype(np.float32)
    feats *= (rng.random((B, R, cfg.feat_dim)) > 0.6)
    c = np.sort(rng.random((B, R, 2, 2)).astype(np.float32), axis=2)  # [.,.,(lo,hi),(x,y)]
    box4 = np.stack([c[:, :, 0, 0], c[:, :, 0, 1], c[:, :, 1, 0], c[:, :, 1, 1]], -1)
    area = ((box4[..., 2] - box4[..., 0]) * (box4[..., 3] - box4[..., 1]))[..., None]
    boxes5 = np.concatenate([box4, area], -1).astype(np.float32)

    # pool of 33 label phrases of 1..T tokens (load_data_v4.py:34-38 maps 33 class ids to phrases)
    prng = np.random.default_rng(SEED0 + 33)
    pool = np.zeros((33, T), np.int32)
    for i in range(33):
        n = int(prng.integers(1, T + 1))
        pool[i, :n] = prng.integers(lo_id, V, n)
    label_ids = pool[rng.integers(0, 33, (B, R))]

    valid = np.arange(R)[None, :] < num_boxes[:, None]
    feats *= valid[..., None]
    boxes5 *= valid[..., None]
    label_ids = label_ids * valid[..., None]

    out = {
        "query_ids": query_ids.astype(np.int32),
        "len_query": len_query,
        "num_boxes": num_boxes,
        "feats": feats.astype(np.float32),
        "label_ids": label_ids.astype(np.int32),
        "query_owner": owner.astype(np.int32),
    }
    if cfg.kind == ZK:
        out["boxes"] = boxes5
        out["segment_ids"] = np.tile(np.array([0] * Lq + [1] * R, np.int32), (B, 1))
        out["labels"] = np.ones(B, np.int32)
    elif cfg.kind == LDS:
        out["boxes"] = boxes5  # fed but unused by the model (run_pretraining_predict_score.py:531)
        out["segment_ids"] = np.zeros((B, Lq), np.int32)
        out["labels"] = np.zeros(B, np.int32)
    else:
        out["boxes"] = boxes5[..., :4].copy()
        out["query_mask"] = (np.arange(Lq)[None, :] < len_query[:, None]).astype(np.int32)
        out["visn_mask"] = valid.astype(np.int32)
        out["label_mask"] = (label_ids > 0).astype(np.int32)
    return out
