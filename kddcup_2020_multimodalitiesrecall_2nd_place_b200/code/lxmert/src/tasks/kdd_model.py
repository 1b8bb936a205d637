"""lxmert/src/tasks/kdd_model.py of the reference, hot-path subset: KDDModel.forward (kdd_model.py:154-214) and the
score extraction of KDD.predict (kdd_model.py:98-112).

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
