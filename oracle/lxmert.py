"""fp32 CPU restatement of the LXMERT scorer.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Pinned: tests/test_oracle_lxmert.py checks this file against golden vectors produced by the reference's OWN code
(oracle/lxmert_ref.py imports /root/reference/code/lxmert/src unmodified; tools/make_golden.py wrote the vectors).

Weights: dict state_dict-key -> fp32 tensor, torch layouts ([out, in]).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

P = "lxrt_encoder.model.bert."


def gelu_erf(x):
    """lxrt/modeling.py:113-119."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def ln(x, w, name):
    return F.layer_norm(x, (x.shape[-1],), w[name + ".weight"], w[name + ".bias"], 1e-12)


def lin(x, w, name):
    return x @ w[name + ".weight"].t() + w[name + ".bias"]


def embeddings(ids, w):
    """BertEmbeddings (modeling.py:269-297): word + position[0..n) + token_type[0], LayerNorm."""
    n = ids.shape[-1]
    x = w[P + "embeddings.word_embeddings.weight"][ids.long()] \
        + w[P + "embeddings.position_embeddings.weight"][:n] \
        + w[P + "embeddings.token_type_embeddings.weight"][0]
    return ln(x, w, P + "embeddings.LayerNorm")


def attention(x, ctx, add_mask, w, prefix, heads):
    """BertAttention (modeling.py:300-352).  add_mask: additive [B,1,1,Sk] or None."""
    B, Sq, H = x.shape
    Sk = ctx.shape[1]
    d = H // heads
    q = lin(x, w, prefix + "query").view(B, Sq, heads, d).permute(0, 2, 1, 3)
    k = lin(ctx, w, prefix + "key").view(B, Sk, heads, d).permute(0, 2, 1, 3)
    v = lin(ctx, w, prefix + "value").view(B, Sk, heads, d).permute(0, 2, 1, 3)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    if add_mask is not None:
        s = s + add_mask
    p = torch.softmax(s, dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B, Sq, H)


def att_block(x, ctx, add_mask, w, att_prefix, out_prefix, heads):
    """BertSelfattLayer / BertCrossattLayer (modeling.py:369-391) = attention + BertAttOutput (355-366)."""
    a = attention(x, ctx, add_mask, w, att_prefix, heads)
    return ln(lin(a, w, out_prefix + "dense") + x, w, out_prefix + "LayerNorm")


def ffn_block(x, w, inter_prefix, out_prefix):
    """BertIntermediate (394-406, erf-GELU) + BertOutput (409-420)."""
    h = gelu_erf(lin(x, w, inter_prefix + "dense"))
    return ln(lin(h, w, out_prefix + "dense") + x, w, out_prefix + "LayerNorm")


def bert_layer(x, add_mask, w, p, heads):
    """BertLayer (modeling.py:423-434)."""
    a = att_block(x, x, add_mask, w, p + "attention.self.", p + "attention.output.", heads)
    return ffn_block(a, w, p + "intermediate.", p + "output.")


def x_layer(lang, lmask, visn, vmask, w, p, heads):
    """LXRTXLayer (modeling.py:444-493): both cross-attentions use the SAME visual_attention weights and the
    pre-update inputs; then per-stream self-attention, then per-stream FFN."""
    l1 = att_block(lang, visn, vmask, w, p + "visual_attention.att.", p + "visual_attention.output.", heads)
    v1 = att_block(visn, lang, lmask, w, p + "visual_attention.att.", p + "visual_attention.output.", heads)
    l2 = att_block(l1, l1, lmask, w, p + "lang_self_att.self.", p + "lang_self_att.output.", heads)
    v2 = att_block(v1, v1, vmask, w, p + "visn_self_att.self.", p + "visn_self_att.output.", heads)
    return ffn_block(l2, w, p + "lang_inter.", p + "lang_output."), ffn_block(v2, w, p + "visn_inter.", p + "visn_output.")


def visual_feat_encoder(feats, boxes, label_emb, w):
    """VisualFeatEncoder (modeling.py:496-533): (LN(fc(f)) + LN(fc(box4)) + LN(fc(conv1x1_8->1(label_emb)))) / 3."""
    v = P + "encoder.visn_fc."
    x = ln(lin(feats, w, v + "visn_fc"), w, v + "visn_layer_norm")
    y = ln(lin(boxes, w, v + "box_fc"), w, v + "box_layer_norm")
    cw = w[v + "label_conv.weight"].reshape(-1)           # [8]
    z = (label_emb * cw[None, None, :, None]).sum(2) + w[v + "label_conv.bias"]   # Conv2d(8,1,1) over the token axis
    z = ln(lin(z, w, v + "label_fc"), w, v + "label_layer_norm")
    return (x + y + z) / 3


@torch.no_grad()
def forward(w, inp, n_l, n_r, n_x, heads=12, with_mlm_head=False):
    """KDDModel.forward (tasks/kdd_model.py:183-214) through LXRTModel.forward (modeling.py:872-927),
    default flags (task_match = task_amsloss = False): logit = logit_fc(pooled)."""
    q, lab = inp["query_ids"], inp["label_ids"]
    lmask = (1.0 - inp["query_mask"].float())[:, None, None, :] * -10000.0       # :890-898
    vmask = (1.0 - inp["visn_mask"].float())[:, None, None, :] * -10000.0        # :904-909
    lang = embeddings(q, w)                                                      # :913
    lemb = embeddings(lab, w)                                                    # :915 (per-sample loop == batched)
    visn = visual_feat_encoder(inp["feats"], inp["boxes"], lemb, w)              # :574
    emb_lang, emb_visn = lang, visn
    for i in range(n_l):
        lang = bert_layer(lang, lmask, w, P + f"encoder.layer.{i}.", heads)      # :577-578
    for i in range(n_r):
        visn = bert_layer(visn, vmask, w, P + f"encoder.r_layers.{i}.", heads)   # :582-583
    for i in range(n_x):
        lang, visn = x_layer(lang, lmask, visn, vmask, w, P + f"encoder.x_layers.{i}.", heads)  # :589-591
    pooled = torch.tanh(lin(lang[:, 0], w, P + "pooler.dense"))                  # BertPooler :596-608
    h = gelu_erf(lin(pooled, w, "logit_fc.0"))                                   # kdd_model.py:167-172
    h = F.layer_norm(h, (h.shape[-1],), w["logit_fc.2.weight"], w["logit_fc.2.bias"], 1e-12)
    logit = lin(h, w, "logit_fc.3")
    probs = torch.softmax(logit, dim=1)                                          # kdd_model.py:102
    x_norm = pooled / pooled.norm(p=2, dim=1, keepdim=True).clamp(min=1e-12)     # :204-205
    return {"probs": probs, "logit": logit, "pooled": pooled, "x_norm": x_norm, "lang": lang, "visn": visn,
            "embedding_output": emb_lang, "visn_embedding": emb_visn, "sequence_output": lang}
