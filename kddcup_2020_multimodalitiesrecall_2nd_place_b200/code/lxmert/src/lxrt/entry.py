"""lxmert/src/lxrt/entry.py of the reference, hot-path subset: LXRTEncoder (entry.py:112-141).

`LXRTEncoder.forward(input_ids, boxes_label_input_ids, segment_ids, input_mask, boxes_label_segment_ids,
boxes_label_input_mask, feats, visual_attention_mask)` with feats = (region features [B,R,2048], boxes [B,R,4])
returns ((lang [B,Lq,768], visn [B,R,768]), pooled [B,768]) as device tensors.  segment ids are all zero in the
reference's drivers (kdd_data.py; modeling.py:287-288) and the label-text masks never reach the encoder
(modeling.py:915 embeds label tokens without a mask), so both are accepted and unused here as well.
"""
from __future__ import annotations

import torch

from .....config import LXMERT
from .... import _runtime as rt


def bind(state_dict, device=0, dtype="fp16", **layers):
    """Replaces KDD.load / load_state_dict (kdd_model.py:131-152): a state_dict keyed like the reference .pth."""
    w = {k[len("module."):] if k.startswith("module.") else k:
         (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in state_dict.items()}
    rt.bind(LXMERT, w, device=device, dtype=dtype, **layers)


def _feeds(input_ids, boxes_label_input_ids, input_mask, feats, boxes, visual_attention_mask):
    input_ids = rt.as_tensor(input_ids, torch.int32)
    feats = rt.as_tensor(feats, torch.float32)
    B, Lq = input_ids.shape
    R = feats.shape[1]
    if input_mask is None:
        input_mask = torch.ones((B, Lq), dtype=torch.int32)                  # modeling.py:884-885
    if visual_attention_mask is None:
        visual_attention_mask = torch.ones((B, R), dtype=torch.int32)
    return {
        "query_ids": input_ids,
        "label_ids": rt.as_tensor(boxes_label_input_ids, torch.int32),
        "feats": feats,
        "boxes": rt.as_tensor(boxes, torch.float32),
        "query_mask": rt.as_tensor(input_mask, torch.int32),
        "visn_mask": rt.as_tensor(visual_attention_mask, torch.int32),
    }, B, Lq, R


class LXRTEncoder(object):
    def __init__(self, args=None, mode="x"):
        self.mode = mode

    @property
    def dim(self):
        return 768

    def forward(self, input_ids, boxes_label_input_ids, segment_ids, input_mask, boxes_label_segment_ids,
                boxes_label_input_mask, feats, visual_attention_mask=None):
        f, b = feats
        feeds, B, Lq, R = _feeds(input_ids, boxes_label_input_ids, input_mask, f, b, visual_attention_mask)
        sc = rt.scorer_for(LXMERT, Lq, R, B)
        out = rt.run(sc, feeds, pooled=True, sequence=True)
        H = sc.cfg.hidden
        lang, visn = [], []
        for lo, chunk in zip(range(0, B, sc.max_batch), out["sequence"]):
            n = min(B, lo + sc.max_batch) - lo
            lang.append(chunk[: n * Lq].view(n, Lq, H))          # rows [0, n*Lq): language stream
            visn.append(chunk[n * Lq: n * (Lq + R)].view(n, R, H))
        return (torch.cat(lang), torch.cat(visn)), out["pooled"]

    __call__ = forward
