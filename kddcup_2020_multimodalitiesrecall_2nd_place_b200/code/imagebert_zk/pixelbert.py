"""imagebert_zk/pixelbert.py of the reference, hot-path subset: BertConfig and BertModel (pixelbert.py:32-107, 150-309).

`BertModel(imgfeat, config, is_training, input_ids, input_mask, token_type_ids, ...)` takes the ALREADY FUSED region
term `imgfeat` [B, R, 768] (label + box + feature, model_triple.py:195) exactly like the reference constructor; the
768->768 `feature_embedding` (pixelbert.py:449-452), the embedding post-processing, the encoder and the pooler run as
sm_100a kernels (mmr_forward with mmr_inputs.region_sum).  Accessors return device tensors.
"""
from __future__ import annotations

import json

import torch

from ...config import ZK
from .. import _runtime as rt


class BertConfig(object):
    """pixelbert.py:32-107 (same fields and defaults; the kernels are built for hidden 768 / 12 heads / 3072)."""

    def __init__(self, vocab_size, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=16,
                 initializer_range=0.02):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.hidden_act = hidden_act
        self.intermediate_size = intermediate_size
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range

    @classmethod
    def from_dict(cls, json_object):
        config = BertConfig(vocab_size=None)
        for key, value in json_object.items():
            config.__dict__[key] = value
        return config

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r") as reader:
            return cls.from_dict(json.loads(reader.read()))

    def to_dict(self):
        return dict(self.__dict__)


def bind(weights, device=0, dtype="fp16", **layers):
    """Replaces saver.restore (evaluate_normal.py:204-212): weights keyed by the checkpoint's variable names."""
    rt.bind(ZK, weights, device=device, dtype=dtype, **layers)


class BertModel(object):
    def __init__(self, imgfeat, config, is_training, input_ids, input_mask=None, token_type_ids=None,
                 use_one_hot_embeddings=False, scope=None, random_sample=True, labels=None):
        if is_training:
            raise NotImplementedError("inference only: dropout / training graphs are outside the scoring hot path")
        input_ids = rt.as_tensor(input_ids, torch.int32)
        B, Lq = input_ids.shape
        imgfeat = rt.as_tensor(imgfeat, torch.float32)
        R = imgfeat.shape[1]
        if input_mask is None:
            input_mask = torch.ones((B, Lq + R), dtype=torch.int32)         # pixelbert.py:189-190
        if token_type_ids is None:
            token_type_ids = torch.zeros((B, Lq + R), dtype=torch.int32)    # pixelbert.py:192-193
        input_mask = rt.as_tensor(input_mask, torch.int32)
        feeds = {
            "query_ids": input_ids,
            "segment_ids": rt.as_tensor(token_type_ids, torch.int32),
            "region_sum": imgfeat,
            "len_query": rt.prefix_lengths(input_mask[:, :Lq], "input_mask[:, :Lq]"),
            "num_boxes": rt.prefix_lengths(input_mask[:, Lq:], "input_mask[:, Lq:]"),
            "labels": rt.as_tensor(labels if labels is not None else torch.ones(B), torch.int32),
        }
        sc = rt.scorer_for(ZK, Lq, R, B)
        out = rt.run(sc, feeds, pooled=True, logits=True, sequence=True, embedding=True, all_layers=True)
        S, H = Lq + R, sc.cfg.hidden
        self.pooled_output = out["pooled"]
        self.sequence_output = torch.cat(out["sequence"]).view(B, S, H)
        self.embedding_output = torch.cat(out["embedding"]).view(B, S, H)
        self.all_encoder_layers = [torch.cat([c[i] for c in out["all_layers"]]).view(B, S, H)
                                   for i in range(sc.cfg.n_layers)]
        self._probs, self._logits = out["probs"], out["logits"]
        self._weights = rt.bound(ZK)["weights"]

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
