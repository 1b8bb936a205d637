"""imagebert_lds/src/run_pretraining_predict_score.py of the reference, hot-path subset (288-394, 479-501).

    probs = bertmodel(bert_config, bert_init_checkpoint, learning_rate, num_train_steps, num_warmup_steps,
                      use_one_hot_embeddings, features, ngpus, is_training=False)          # [B, 2]

`features` keeps the reference keys (input_ids, segment_ids, boxes, features, labelfeat, next_sentence_labels,
query_id, product_id; run_pretraining_predict_score.py:526-548).  `bert_init_checkpoint` is a dict of weights keyed
by the checkpoint's variable names (or None when `pixelmodel.bind` was called); optimiser arguments are accepted
and unused at inference, as in the reference's `is_training=False` branch.
"""
from __future__ import annotations

import torch

from ....config import LDS
from ... import _runtime as rt
from . import pixelmodel


def get_next_sentence_output(bert_config, input_tensor, labels):
    """(loss, per_example_loss, log_probs, probs) of the 2-way head on pooled rows (:479-501)."""
    w = rt.bound(LDS)["weights"]
    probs, logits = rt.linear_head(input_tensor, w["cls/seq_relationship/output_weights"],
                                   w["cls/seq_relationship/output_bias"])
    labels = rt.as_tensor(labels, torch.int64).view(-1)
    loss, per_example, log_probs = rt.cross_entropy(logits, labels)
    return loss, per_example, log_probs, probs


def bertmodel(bert_config, bert_init_checkpoint, learning_rate, num_train_steps, num_warmup_steps,
              use_one_hot_embeddings, features, ngpus, is_training):
    if is_training:
        raise NotImplementedError("inference only (the reference scores with is_training=False, :566-576)")
    if isinstance(bert_init_checkpoint, dict):
        if rt._BOUND.get(LDS, {}).get("weights") is not bert_init_checkpoint:
            pixelmodel.bind(bert_init_checkpoint)
    input_ids = rt.as_tensor(features["input_ids"], torch.int32)
    feats = rt.as_tensor(features["features"], torch.float32)
    B, Lq = input_ids.shape
    R = feats.shape[1]
    feeds = {
        "query_ids": input_ids,
        "segment_ids": rt.as_tensor(features["segment_ids"], torch.int32),
        "label_ids": rt.as_tensor(features["labelfeat"], torch.int32),
        "feats": feats,
    }
    sc = rt.scorer_for(LDS, Lq, R, B)
    return rt.run(sc, feeds)["probs"]
