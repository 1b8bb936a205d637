"""imagebert_zk/model_triple.py of the reference, hot-path subset (model_triple.py:56-106, 162-214).

    model_triple.bind(weights)                       # was: saver.restore(sess, FLAGS.pretrained_model_path)
    loss, probs, loss_list = model_triple.model_attention_channel_e(
        num_boxes, np_boxes_5, np_images_features, np_idx_class_labels, np_len_class_labels, np_idx_query_,
        len_query_, labels, segment_ids, label_query, weight_label_query, is_training=False)

Arguments keep the reference's order, shapes and dtypes (feed shapes: evaluate_normal.py:141-152); arrays may be
numpy or torch, on the host or already on the GPU.  The score of a pair is probs[:, 1] (evaluate_normal.py:240).
"""
from __future__ import annotations

import torch

from ...config import ZK
from .. import _runtime as rt
from . import pixelbert

bert_config = None   # the reference module builds it from ../user_data/bert_config.json at import (model_triple.py:24)


def bind(weights, device=0, dtype="fp16", **layers):
    rt.bind(ZK, weights, device=device, dtype=dtype, **layers)


def amsoftmax_loss(y_true, y_pred):
    """(cross_ent [B], probs [B,2]) of the AM-softmax head; y_true one-hot [B,2], y_pred = pooled output [B,768]."""
    y_true = rt.as_tensor(y_true, torch.float32)
    labels = y_true.argmax(-1)
    w = rt.bound(ZK)["weights"]
    probs, logits = rt.am_softmax_head(y_pred, w["cls/seq_relationship/am_kernel"], labels)
    _, per_example, _ = rt.cross_entropy(logits, labels)
    return per_example, probs


def get_next_sentence_output_am(input_tensor, labels):
    labels = rt.as_tensor(labels, torch.int64).view(-1)
    one_hot = torch.nn.functional.one_hot(labels, 2).float()
    loss, probs = amsoftmax_loss(one_hot, input_tensor)
    return loss.mean(), probs


def image_bert(np_images_features, np_idx_query_, is_training, input_mask, segment_ids):
    """model_triple.py:108-121: the BertModel over [query ; fused region term]."""
    return pixelbert.BertModel(imgfeat=np_images_features, config=bert_config, is_training=is_training,
                               input_ids=np_idx_query_, input_mask=input_mask, token_type_ids=segment_ids)


def model_attention_channel_e(num_boxes, np_boxes_5, np_images_features, np_idx_class_labels, np_len_class_labels,
                              np_idx_query_, len_query_, labels, segment_ids, label_query, weight_label_query,
                              is_training=True, reuse=None):
    """Whole zk scorer in one fused device pass (label conv as table gathers, box FC, feature conv + ReLU, sum,
    feature_embedding, embeddings + LayerNorm, 12 encoder layers, pooler, AM-softmax).  Returns
    (loss, probs [B,2], [loss]) like the reference; np_len_class_labels / label_query / weight_label_query are
    accepted and unused, as in the reference graph (model_triple.py:162-214)."""
    if is_training:
        raise NotImplementedError("inference only (the reference scores with is_training_tensor: False, "
                                  "evaluate_normal.py:236)")
    query_ids = rt.as_tensor(np_idx_query_, torch.int32)
    feats = rt.as_tensor(np_images_features, torch.float32)
    B, Lq = query_ids.shape
    R = feats.shape[1]
    labels_t = rt.as_tensor(labels, torch.int32).view(-1)
    feeds = {
        "query_ids": query_ids,
        "segment_ids": rt.as_tensor(segment_ids, torch.int32),
        "label_ids": rt.as_tensor(np_idx_class_labels, torch.int32),
        "feats": feats,
        "boxes": rt.as_tensor(np_boxes_5, torch.float32),
        "len_query": rt.as_tensor(len_query_, torch.int32).view(-1),
        "num_boxes": rt.as_tensor(num_boxes, torch.int32).view(-1),
        "labels": labels_t,
    }
    sc = rt.scorer_for(ZK, Lq, R, B)
    out = rt.run(sc, feeds, logits=True)
    loss, _, _ = rt.cross_entropy(out["logits"], labels_t)
    return loss, out["probs"], [loss]
