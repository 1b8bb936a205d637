"""imagebert_lds/src/pixelmodel.py of the reference, hot-path subset: BertModel (pixelmodel.py:145-270).

`BertModel(imgfeat, config, is_training, input_ids, label_ids, input_mask, token_type_ids, ...)`: imgfeat is the RAW
2048-d region features [B, R, 2048], label_ids [B, R, 8]; the reference ignores input_mask (all-ones attention mask,
pixelmodel.py:189-192) and so does this class.
"""
from __future__ import annotations

import torch

from ....config import LDS
from ... import _runtime as rt
from ...imagebert_zk.pixelbert import BertConfig  # noqa: F401  (same class in both reference trees)


def bind(weights, device=0, dtype="fp16", **layers):
    rt.bind(LDS, weights, device=device, dtype=dtype, **layers)


class BertModel(object):
    def __init__(self, imgfeat, config, is_training, input_ids, label_ids, input_mask=None, token_type_ids=None,
                 use_one_hot_embeddings=False, scope=None, random_sample=True):
        if is_training:
            raise NotImplementedError("inference only: dropout / training graphs are outside the scoring hot path")
        input_ids = rt.as_tensor(input_ids, torch.int32)
        B, Lq = input_ids.shape
        imgfeat = rt.as_tensor(imgfeat, torch.float32)
        R = imgfeat.shape[1]
        if token_type_ids is None:
            token_type_ids = torch.zeros((B, Lq), dtype=torch.int32)
        feeds = {
            "query_ids": input_ids,
            "segment_ids": rt.as_tensor(token_type_ids, torch.int32),
            "label_ids": rt.as_tensor(label_ids, torch.int32),
            "feats": imgfeat,
        }
        sc = rt.scorer_for(LDS, Lq, R, B)
        out = rt.run(sc, feeds, pooled=True, logits=True, sequence=True, embedding=True, all_layers=True)
        S, H = Lq + 2 * R, sc.cfg.hidden
        self.pooled_output = out["pooled"]
        self.sequence_output = torch.cat(out["sequence"]).view(B, S, H)
        self.embedding_output = torch.cat(out["embedding"]).view(B, S, H)
        self.all_encoder_layers = [torch.cat([c[i] for c in out["all_layers"]]).view(B, S, H)
                                   for i in range(sc.cfg.n_layers)]
        self._probs, self._logits = out["probs"], out["logits"]
        self._weights = rt.bound(LDS)["weights"]

    def get_pooled_output(self):
        return self.pooled_output

    def get_sequence_output(self):
        return self.sequence_output

    def get_all_encoder_layers(self):
        return self.all_encoder_layers

    def get_embedding_output(self):
        return self.embedding_output

    def get_embedding_table(self):
        return self._weights["bert/embeddings/word_embeddings"]
