"""fp32 CPU restatement of the two TF-1 ImageBert scorers.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Pinning: TensorFlow 1.12 / Python 2 cannot run here and the reference ships neither weights nor activations for
these graphs.  What pins this file is the reference's OWN model code (imagebert_zk/pixelbert.py + model_triple.py,
imagebert_lds/src/pixelmodel.py + get_next_sentence_output) executed unmodified on tools/tf1_shim.py — an eager
stand-in for the TensorFlow ops it calls — on seeded synthetic weights / inputs: tests/golden/{zk,lds}_ref_shim_*.npz
(tools/make_golden.py --tf-shim), compared in tests/test_oracle.py to 1e-5 on probs, pooled output and per-token
statistics of the embedding and final layers, 2 and 12 layers, reference-initialiser and trained-like weights.  That
pins the WIRING (ops, order, variable names, shapes, masks, constants, quirks) to the reference's source.  The
arithmetic of the individual TensorFlow ops (conv2d SAME / default ReLU, fully_connected, dense, layer_norm eps,
l2_normalize, dropout at inference) is restated in the shim from the TF 1.12 documentation: PARITY UNPINNED at that
level, and only there.

Weights: dict name -> fp32 tensor with the reference's TF variable names and layouts (kernels are [in, out]).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def gelu_tanh(x):
    """imagebert_zk/pixelbert.py:315-328."""
    return x * 0.5 * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def layer_norm(x, gamma, beta):
    """tf.contrib.layers.layer_norm(begin_norm_axis=-1), pixelbert.py:414-417: biased variance, eps 1e-12."""
    return F.layer_norm(x, (x.shape[-1],), gamma, beta, 1e-12)


def dense(x, w, name):
    """tf.layers.dense: x @ kernel[in,out] + bias."""
    return x @ w[name + "/kernel"] + w[name + "/bias"]


def attention_layer(x, key_mask, w, prefix, heads):
    """pixelbert.py:658-852 with from_tensor == to_tensor.  x [B,S,H]; key_mask float [B,S] (1 = attend)."""
    B, S, H = x.shape
    d = H // heads
    q = dense(x, w, prefix + "/query").view(B, S, heads, d).transpose(1, 2)   # :767-795
    k = dense(x, w, prefix + "/key").view(B, S, heads, d).transpose(1, 2)
    v = dense(x, w, prefix + "/value").view(B, S, heads, d).transpose(1, 2)
    scores = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(float(d)))          # :802-804
    if key_mask is not None:
        scores = scores + (1.0 - key_mask)[:, None, None, :] * -10000.0       # :806-817
    probs = torch.softmax(scores, dim=-1)                                      # :821
    ctx = probs @ v                                                            # :836
    return ctx.transpose(1, 2).reshape(B, S, H)                                # :839-850


def transformer_model(x, key_mask, w, n_layers, heads):
    """pixelbert.py:855-995 (identical core in imagebert_lds/src/pixelmodel.py:836-974)."""
    layers = []
    for i in range(n_layers):
        p = f"bert/encoder/layer_{i}"
        ctx = attention_layer(x, key_mask, w, p + "/attention/self", heads)
        att = dense(ctx, w, p + "/attention/output/dense")                                      # :960-964
        att = layer_norm(att + x, w[p + "/attention/output/LayerNorm/gamma"],
                         w[p + "/attention/output/LayerNorm/beta"])                               # :966
        inter = gelu_tanh(dense(att, w, p + "/intermediate/dense"))                              # :969-974
        out = dense(inter, w, p + "/output/dense")                                              # :977-981
        x = layer_norm(out + att, w[p + "/output/LayerNorm/gamma"], w[p + "/output/LayerNorm/beta"])  # :983
        layers.append(x)
    return layers


def pooler(seq, w):
    """pixelbert.py:258-266: tanh(dense(first token))."""
    return torch.tanh(dense(seq[:, 0], w, "bert/pooler/dense"))


# ---------------------------------------------------------------------------------------------- zk
def zk_label_term(label_ids, w):
    """model_triple.py:178-190: gather E[ids] -> slim.conv2d(768,[1,8]) SAME + bias + default ReLU -> mean over
    the 8 positions.  SAME padding of the even 8-tap kernel: 3 left, 4 right.  [PAD]=0 embeds to E[0]."""
    E = w["bert/embeddings/word_embeddings"]
    lab = E[label_ids.long()]                      # [B,R,8,H]
    B, R, T, H = lab.shape
    Wc = w["kdd_conv1/weights"][0]                 # [8, Hin, Hout]
    taps = Wc.shape[0]
    pad_l = (taps - 1) // 2                        # 3
    pad_r = taps - 1 - pad_l                       # 4
    labp = F.pad(lab, (0, 0, pad_l, pad_r))        # pad the token axis
    out = torch.zeros(B, R, T, Wc.shape[2], dtype=lab.dtype)
    for k in range(taps):
        out = out + labp[:, :, k:k + T, :] @ Wc[k]
    out = torch.relu(out + w["kdd_conv1/biases"])
    return out.mean(dim=2)


def zk_region_tokens(feats, boxes5, label_ids, w):
    """model_triple.py:189-195 + pixelbert.feature_embedding (pixelbert.py:449-452, called at :186)."""
    label = zk_label_term(label_ids, w)
    box = boxes5 @ w["kdd_dense1/weights"] + w["kdd_dense1/biases"]                       # :191 (no activation)
    feat = torch.relu(feats @ w["kdd_conv2/weights"][0, 0] + w["kdd_conv2/biases"])       # :192-194 (default ReLU)
    region = label + box + feat                                                           # :195
    return region @ w["kdd_featureemb/fully_connected/weights"] + w["kdd_featureemb/fully_connected/biases"]


def zk_embeddings(query_ids, segment_ids, region, w):
    """pixelbert.py:493-538 + 541-621: word gather, concat [text; region], + type, + position
    [0..Lq-1] + [Lq]*R (generalised from the hard-coded range(20)+[20]*10 at :614), LayerNorm."""
    E = w["bert/embeddings/word_embeddings"]
    B, Lq = query_ids.shape
    R = region.shape[1]
    x = torch.cat([E[query_ids.long()], region], dim=1)
    x = x + w["bert/embeddings/token_type_embeddings"][segment_ids.long()]
    pos_idx = torch.tensor(list(range(Lq)) + [Lq] * R, dtype=torch.long)
    x = x + w["bert/embeddings/position_embeddings"][pos_idx]
    return layer_norm(x, w["bert/embeddings/LayerNorm/gamma"], w["bert/embeddings/LayerNorm/beta"])


def amsoftmax_probs(pooled, labels, w):
    """model_triple.py:56-86: scale 30, margin 0.35 applied to the fed label's cosine when it exceeds 0.35."""
    x = pooled * torch.rsqrt(torch.clamp((pooled * pooled).sum(1, keepdim=True), min=1e-12))   # l2_normalize dim=1
    k = w["cls/seq_relationship/am_kernel"]
    kn = k * torch.rsqrt(torch.clamp((k * k).sum(0, keepdim=True), min=1e-10))                 # :64
    c = torch.clamp(x @ kn, -1.0, 1.0)
    y = F.one_hot(labels.long(), 2).to(c.dtype)
    g = (c * y).sum(1, keepdim=True)
    added = (g > 0.35).to(c.dtype) * 0.35
    logits = (c - y * added) * 30.0
    return torch.softmax(logits, dim=-1)


@torch.no_grad()
def zk_forward(w, inp, n_layers, heads=12):
    """model_triple.model_attention_channel_e (model_triple.py:162-214), inference.  Returns dict."""
    query_ids, seg = inp["query_ids"], inp["segment_ids"]
    B, Lq = query_ids.shape
    R = inp["feats"].shape[1]
    region = zk_region_tokens(inp["feats"], inp["boxes"], inp["label_ids"], w)
    x0 = zk_embeddings(query_ids, seg, region, w)
    qmask = torch.arange(Lq)[None, :] < inp["len_query"].long()[:, None]         # :198 tf.sequence_mask
    bmask = torch.arange(R)[None, :] < inp["num_boxes"].long()[:, None]          # :199
    key_mask = torch.cat([qmask, bmask], dim=1).to(x0.dtype)
    layers = transformer_model(x0, key_mask, w, n_layers, heads)
    pooled = pooler(layers[-1], w)
    probs = amsoftmax_probs(pooled, inp["labels"], w)
    return {"probs": probs, "pooled": pooled, "embedding_output": x0, "sequence_output": layers[-1],
            "all_encoder_layers": layers, "region": region}


# ---------------------------------------------------------------------------------------------- lds
def lds_label_term(label_ids, w):
    """pixelmodel.py:489-498 (reshape4D=True): gather [B*R*8, H] -> reshape(-1, 8) -> @ [8,1] -> [B,R,H].
    Output dim j mixes 8 CONSECUTIVE hidden dims of token floor(8j/H) (a reference quirk, kept)."""
    E = w["bert/embeddings/word_embeddings"]
    B, R, T = label_ids.shape
    g = E[label_ids.long().reshape(-1)]                       # [B*R*T, H]
    H = g.shape[1]
    g = g.reshape(-1, T)                                      # row-major regroup
    out = g @ w["bert/embeddings/word_embeddings_labelembedding"]   # [B*R*H, 1]
    return out.squeeze(-1).reshape(B, R, H)


@torch.no_grad()
def lds_forward(w, inp, n_layers, heads=12):
    """pixelmodel.BertModel (pixelmodel.py:145-270) + get_next_sentence_output
    (run_pretraining_predict_score.py:479-501), is_training=False."""
    E = w["bert/embeddings/word_embeddings"]
    query_ids, seg = inp["query_ids"], inp["segment_ids"]
    B, Lq = query_ids.shape
    region = inp["feats"] @ w["featureemb/fully_connected/weights"] + w["featureemb/fully_connected/biases"]  # :439-442
    label = lds_label_term(inp["label_ids"], w)
    text = E[query_ids.long()] + w["bert/embeddings/token_type_embeddings"][seg.long()] \
        + w["bert/embeddings/position_embeddings"][:Lq]                                   # :560-597
    text = layer_norm(text, w["bert/embeddings/LayerNorm/gamma"], w["bert/embeddings/LayerNorm/beta"])  # :600
    x0 = torch.cat([text, region, label], dim=1)                                          # :601
    layers = transformer_model(x0, None, w, n_layers, heads)                              # all-ones mask :189-192
    pooled = pooler(layers[-1], w)
    logits = pooled @ w["cls/seq_relationship/output_weights"].t() + w["cls/seq_relationship/output_bias"]
    probs = torch.softmax(logits, dim=-1)
    return {"probs": probs, "pooled": pooled, "embedding_output": x0, "sequence_output": layers[-1],
            "all_encoder_layers": layers}


def to_torch(d):
    import numpy as np
    return {k: (torch.from_numpy(np.ascontiguousarray(v)) if not torch.is_tensor(v) else v) for k, v in d.items()}
