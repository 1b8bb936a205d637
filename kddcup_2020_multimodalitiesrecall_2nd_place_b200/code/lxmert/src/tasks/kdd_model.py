"""lxmert/src/tasks/kdd_model.py of the reference, hot-path subset: KDDModel.forward (kdd_model.py:154-214) and
KDD.predict (kdd_model.py:46-129), the scoring loop over a data split.

    model = KDDModel(); model.load_state_dict(state_dict)
    x_norm, lang_prediction_scores, logit = model(input_ids, boxes_label_input_ids, segment_ids, input_mask,
                                                  boxes_label_segment_ids, boxes_label_input_mask, feats, boxes,
                                                  visual_attention_mask)

`lang_prediction_scores` (the masked-LM head over all query tokens, 1.08 GFLOP/pair) is computed and discarded by
the reference at inference (kdd_model.py:201-202 vs 98-112); it is returned as None.  The match branch is
`logit_fc` (args.task_amsloss is False in the shipped run scripts).
"""
from __future__ import annotations

import torch

from .....config import LXMERT
from .... import _runtime as rt
from ..lxrt import entry

MAX_LENGTH = 23
MAX_BOX_NUM = 10
MAX_LABLETEXT_LENGTH = 8


class KDDModel(object):
    def __init__(self):
        self.lxrt_encoder = entry.LXRTEncoder(mode="lx")

    def load_state_dict(self, state_dict, strict=False, device=0, dtype="fp16", **layers):
        entry.bind(state_dict, device=device, dtype=dtype, **layers)

    def eval(self):
        return self

    def forward(self, input_ids, boxes_label_input_ids, segment_ids, input_mask, boxes_label_segment_ids,
                boxes_label_input_mask, feats, boxes, visual_attention_mask):
        feeds, B, Lq, R = entry._feeds(input_ids, boxes_label_input_ids, input_mask, feats, boxes,
                                       visual_attention_mask)
        sc = rt.scorer_for(LXMERT, Lq, R, B)
        out = rt.run(sc, feeds, pooled=True, logits=True)
        pooled = out["pooled"]
        x_norm = pooled / pooled.norm(p=2, dim=1, keepdim=True).clamp(min=1e-12)     # kdd_model.py:204-205
        self._probs = out["probs"]
        return x_norm, None, out["logits"]

    __call__ = forward

    def rank_scores(self, logit=None):
        """Softmax(dim=1)(logit)[:, -1], the per-pair score KDD.predict ranks by (kdd_model.py:102-112)."""
        if logit is None:
            return self._probs[:, 1]
        return torch.softmax(logit, dim=1)[:, -1]


class KDD(object):
    """The predict side of the reference's KDD class (kdd_model.py:26-152): load a `.pth`, score a split, optionally
    save `<result>/<mod>_score_lxmert.csv`.

        kdd = KDD(data_dir, result_dir, vocab_file, label_file); kdd.load(".../BEST"); kdd.predict("testB", save=True)
    """

    def __init__(self, data_dir: str, result_dir: str, vocab_file: str, label_file: str, device: int = 0,
                 dtype: str = "fp16", precision: str = "fast", batch_size: int = 256):
        from ..... import records, tokenizer
        self.data_dir, self.result_dir = data_dir, result_dir
        self.device, self.dtype, self.precision, self.batch_size = device, dtype, precision, batch_size
        self.tokenizer = tokenizer.FullTokenizer(vocab_file=vocab_file, max_input_chars_per_word=100)   # lxrt/tokenization.py:298
        self.label_map = records.load_label_map(label_file)
        self.scorer = None

    def load(self, path: str, **layers):
        """kdd_model.py:131-152: torch.load(path + ".pth"), DataParallel prefixes stripped, non-strict."""
        from ..... import checkpoints
        from .....config import ModelConfig
        from .....scorer import MatchScorer
        w = checkpoints.load_pth(path)
        depth = rt._depth(LXMERT, w, layers)
        vocab = w["lxrt_encoder.model.bert.embeddings.word_embeddings.weight"].shape[0]
        cfg = ModelConfig(LXMERT, lq=MAX_LENGTH, nbox=MAX_BOX_NUM, vocab=int(vocab), **depth)
        if self.scorer is not None:
            self.scorer.close()
        self.scorer = MatchScorer(cfg, w, device=self.device, dtype=self.dtype, max_batch=self.batch_size,
                                  precision=self.precision)
        return self

    def predict(self, mod: str = "valid", save: bool = False, rank: int = 0, world: int = 1):
        """(match_pred, match_label, rank_score_pred) as kdd_model.py:46-129 returns them: argmax of the 2-way softmax
        per pair, the label column (the test loaders feed zeros: kdd_data.py), and {query id: [(product id, score)]}
        with score = Softmax(1)(logit)[:, -1]."""
        import collections
        import os

        import numpy as np

        from ..... import drivers
        if self.scorer is None:
            raise RuntimeError("KDD.predict: call load(path) first")
        lines = drivers.read_tsv_lines(os.path.join(self.data_dir, mod, f"{mod}.tsv"))
        res = drivers.score_tsv(self.scorer, self.tokenizer, self.label_map, lines, rank=rank, world=world)
        rank_score_pred = collections.defaultdict(list)
        for q, p, s in zip(res["query_id"].tolist(), res["product_id"].tolist(), res["score"].tolist()):
            rank_score_pred[q].append((p, s))
        match_pred = (res["score"] > 0.5).astype(np.int64).tolist()        # argmax over [1 - s, s]
        match_label = [0] * len(lines)
        if save and rank == 0:
            drivers.write_scores(os.path.join(self.result_dir, f"{mod}_score_lxmert.csv"), res, lxmert_csv=True)
        return match_pred, match_label, rank_score_pred
