"""Model configuration of the three scorers (one dataclass instead of the reference's scattered constants).

Reference sources: code/user_data/bert_config.json (hidden 768, 12 layers, 12 heads, intermediate 3072,
max_position 512, type_vocab 2, vocab 21128); hard-coded MAX_LENGTH / MAX_BOX_NUM / label length at
imagebert_zk/load_data_v4.py:28-29, lxmert/src/tasks/kdd_model.py:21-23; lxmert layer counts at
lxmert/src/param.py:79-81.  The reference hard-codes (20 query tokens, 10 boxes); here (lq, nbox) are
parameters so the BASELINE configs (32 x 36) run through the same code.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass

ZK, LDS, LXMERT = "imagebert_zk", "imagebert_lds", "lxmert"
KIND_CODE = {ZK: 0, LDS: 1, LXMERT: 2}


@dataclass(frozen=True)
class ModelConfig:
    kind: str
    n_layers: int = 12       # single-stream depth (zk / lds); LXMERT language layers (param.py:79 -> 9)
    n_r_layers: int = 0      # LXMERT relational layers (param.py:81 -> 5)
    n_x_layers: int = 0      # LXMERT cross-modality layers (param.py:80 -> 5)
    lq: int = 20             # query tokens
    nbox: int = 10           # region slots
    hidden: int = 768
    heads: int = 12
    intermediate: int = 3072
    vocab: int = 21128
    max_pos: int = 512
    type_vocab: int = 2
    feat_dim: int = 2048
    label_len: int = 8

    @property
    def seq_len(self) -> int:
        """Tokens per pair in the (first) encoder stream."""
        if self.kind == ZK:
            return self.lq + self.nbox
        if self.kind == LDS:
            return self.lq + 2 * self.nbox
        return self.lq

    @property
    def box_dim(self) -> int:
        return 4 if self.kind == LXMERT else 5

    def to_dict(self):
        return asdict(self)


def native(kind: str) -> ModelConfig:
    """The shapes the reference drivers actually run."""
    if kind == ZK:
        return ModelConfig(ZK, n_layers=12, lq=20, nbox=10)
    if kind == LDS:
        return ModelConfig(LDS, n_layers=12, lq=20, nbox=10)
    if kind == LXMERT:
        return ModelConfig(LXMERT, n_layers=9, n_r_layers=5, n_x_layers=5, lq=23, nbox=10)
    raise ValueError(kind)


def baseline_cfg2(kind: str = ZK) -> ModelConfig:
    """BASELINE.json configs[1]/[2]: 32 query tokens x 36 regions x 2048-d."""
    if kind == LXMERT:
        return ModelConfig(LXMERT, n_layers=9, n_r_layers=5, n_x_layers=5, lq=32, nbox=36)
    return ModelConfig(kind, n_layers=12, lq=32, nbox=36)


def flops_per_pair(cfg: ModelConfig) -> int:
    """Algorithmic FLOPs per (query, product) pair, multiply-add = 2 (SURVEY.md section 8d / BASELINE.md section 3).

    Core figure only: region projection + encoder + pooler/head; variant extras (zk label conv, featureemb)
    and LXMERT's dead MLM head are excluded, exactly as BASELINE.md counts them.
    """
    H, I = cfg.hidden, cfg.intermediate
    g = 2 * (4 * H * H + 2 * H * I)

    def att(sq, sk):
        return 4 * sq * sk * H

    if cfg.kind in (ZK, LDS):
        S = cfg.seq_len
        return cfg.n_layers * (S * g + att(S, S)) + 2 * cfg.nbox * cfg.feat_dim * H + 2 * H * H + 4 * H
    L, R = cfg.lq, cfg.nbox
    lang = cfg.n_layers * (L * g + att(L, L))
    visn = cfg.n_r_layers * (R * g + att(R, R))
    # x-layer: shared cross-attention both ways (QKV + out-proj on both streams), self-attention + FFN per stream
    proj = 2 * 4 * H * H          # q,k,v,o projections per token
    ffn = 2 * 2 * H * I
    x = (L + R) * proj + att(L, R) + att(R, L) + (L + R) * proj + att(L, L) + att(R, R) + (L + R) * ffn
    vis_embed = 2 * R * cfg.feat_dim * H + 2 * R * 4 * H + 2 * R * H * H + 2 * R * 8 * H
    head = 2 * H * 2 * H + 2 * 2 * H * 2
    return lang + visn + cfg.n_x_layers * x + vis_embed + 2 * H * H + head
