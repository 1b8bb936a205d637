"""The reference-named entry points (kddcup_..._b200/code/**, same module paths / names / argument order as the
reference's /code tree) against the fp32 oracle, called the way the reference drivers call them."""
import numpy as np
import pytest
import torch

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import synth
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LDS, LXMERT, ZK, ModelConfig

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _oracle(cfg, w, inp):
    from oracle import imagebert, lxmert
    wt, it = imagebert.to_torch(w), imagebert.to_torch(inp)
    if cfg.kind == ZK:
        return imagebert.zk_forward(wt, it, cfg.n_layers), wt, it
    if cfg.kind == LDS:
        return imagebert.lds_forward(wt, it, cfg.n_layers), wt, it
    return lxmert.forward(wt, it, cfg.n_layers, cfg.n_r_layers, cfg.n_x_layers), wt, it


def test_zk_model_attention_channel_e_and_bertmodel():
    """evaluate_normal.py:222-240: probs = model_attention_channel_e(feeds...)[1]; score = probs[:, 1]."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code.imagebert_zk import model_triple, pixelbert
    from oracle import imagebert
    cfg = ModelConfig(ZK, n_layers=2, lq=20, nbox=10, vocab=2000)      # the reference's native 20 + 10 shape
    w = synth.make_weights(cfg, seed=11)
    inp = synth.make_inputs(cfg, 6, seed=11, n_queries=2)
    ref, wt, it = _oracle(cfg, w, inp)
    model_triple.bind(w)
    B = 6
    loss, probs, loss_list = model_triple.model_attention_channel_e(
        inp["num_boxes"], inp["boxes"], inp["feats"], inp["label_ids"], np.zeros((B, cfg.nbox), np.int32),
        inp["query_ids"], inp["len_query"], inp["labels"].astype(np.int64), inp["segment_ids"],
        np.zeros((B, 18), np.int64), np.zeros((B, 18), np.float32), is_training=False)
    assert (probs.cpu() - ref["probs"]).abs().max().item() <= TOL
    ref_loss = -torch.log(ref["probs"][torch.arange(B), it["labels"].long()]).mean()
    assert abs(loss.item() - ref_loss.item()) < 5e-3 and len(loss_list) == 1
    with pytest.raises(NotImplementedError):
        model_triple.model_attention_channel_e(*([None] * 11), is_training=True)
    # pixelbert.BertModel on the pre-fused region term (model_triple.py:195 -> image_bert -> BertModel)
    region_sum = (imagebert.zk_label_term(it["label_ids"], wt)
                  + it["boxes"] @ wt["kdd_dense1/weights"] + wt["kdd_dense1/biases"]
                  + torch.relu(it["feats"] @ wt["kdd_conv2/weights"][0, 0] + wt["kdd_conv2/biases"]))
    mask = np.concatenate([np.arange(cfg.lq)[None] < inp["len_query"][:, None],
                           np.arange(cfg.nbox)[None] < inp["num_boxes"][:, None]], 1).astype(np.int32)
    model = model_triple.image_bert(region_sum.numpy(), inp["query_ids"], False, mask, inp["segment_ids"])
    assert (model.get_pooled_output().cpu() - ref["pooled"]).abs().max().item() < 2e-2
    assert (model.get_sequence_output().cpu() - ref["sequence_output"]).abs().max().item() < 2e-2
    assert (model.get_embedding_output().cpu() - ref["embedding_output"]).abs().max().item() < 1e-2
    layers = model.get_all_encoder_layers()
    assert len(layers) == 2
    for got, want in zip(layers, ref["all_encoder_layers"]):
        assert (got.cpu() - want).abs().max().item() < 2e-2
    loss2, probs2 = model_triple.get_next_sentence_output_am(model.get_pooled_output(), inp["labels"])
    assert (probs2.cpu() - ref["probs"]).abs().max().item() <= TOL
    with pytest.raises(ValueError, match="prefix"):
        bad = mask.copy()
        bad[0, 0] = 0
        pixelbert.BertModel(region_sum.numpy(), None, False, inp["query_ids"], bad, inp["segment_ids"])
    pixelbert.rt.release()


def test_lds_bertmodel_and_head():
    """run_pretraining_predict_score.py:566-576: probs = bertmodel(..., features, ngpus, is_training=False)."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code.imagebert_lds.src import (
        pixelmodel, run_pretraining_predict_score as rp)
    cfg = ModelConfig(LDS, n_layers=2, lq=20, nbox=10, vocab=2000)
    w = synth.make_weights(cfg, seed=12)
    inp = synth.make_inputs(cfg, 5, seed=12)
    ref, _, _ = _oracle(cfg, w, inp)
    features = {"input_ids": inp["query_ids"], "segment_ids": inp["segment_ids"], "boxes": inp["boxes"],
                "features": inp["feats"], "labelfeat": inp["label_ids"],
                "next_sentence_labels": inp["labels"], "query_id": np.arange(5), "product_id": np.arange(5)}
    probs = rp.bertmodel(None, w, 1e-5, 0, 0, False, features, 1, is_training=False)
    assert (probs.cpu() - ref["probs"]).abs().max().item() <= TOL
    m = pixelmodel.BertModel(inp["feats"], None, False, inp["query_ids"], inp["label_ids"], None,
                             inp["segment_ids"])
    assert (m.get_pooled_output().cpu() - ref["pooled"]).abs().max().item() < 2e-2
    assert (m.get_sequence_output().cpu() - ref["sequence_output"]).abs().max().item() < 2e-2
    loss, per_ex, log_probs, probs2 = rp.get_next_sentence_output(None, m.get_pooled_output(), inp["labels"])
    assert (probs2.cpu() - ref["probs"]).abs().max().item() <= TOL
    assert torch.allclose(log_probs.exp().cpu(), probs2.cpu(), atol=1e-5) and per_ex.shape == (5,)
    pixelmodel.rt.release()


def test_lxmert_kddmodel_forward():
    """kdd_model.py:98-112: x_norm, _, logit = model(...); score = Softmax(dim=1)(logit)[:, -1]."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code.lxmert.src.tasks.kdd_model import KDDModel
    cfg = ModelConfig(LXMERT, n_layers=2, n_r_layers=1, n_x_layers=2, lq=23, nbox=10, vocab=2000)
    w = synth.make_weights(cfg, seed=13)
    inp = synth.make_inputs(cfg, 5, seed=13)
    ref, _, _ = _oracle(cfg, w, inp)
    model = KDDModel()
    model.load_state_dict({("module." + k): torch.from_numpy(v) for k, v in w.items()})   # DataParallel-style keys
    B, R = 5, cfg.nbox
    x_norm, mlm, logit = model(torch.from_numpy(inp["query_ids"]).long(), torch.from_numpy(inp["label_ids"]).long(),
                               torch.zeros(B, cfg.lq, dtype=torch.long), torch.from_numpy(inp["query_mask"]).long(),
                               torch.zeros(B, R, 8, dtype=torch.long), torch.from_numpy(inp["label_mask"]).long(),
                               torch.from_numpy(inp["feats"]), torch.from_numpy(inp["boxes"]),
                               torch.from_numpy(inp["visn_mask"]).long())
    assert mlm is None
    score = torch.softmax(logit, dim=1)[:, -1]
    assert (score.cpu() - ref["probs"][:, 1]).abs().max().item() <= TOL
    assert (model.rank_scores().cpu() - ref["probs"][:, 1]).abs().max().item() <= TOL
    assert (x_norm.cpu() - ref["x_norm"]).abs().max().item() < 2e-3
    (lang, visn), pooled = model.lxrt_encoder(inp["query_ids"], inp["label_ids"], None, inp["query_mask"], None, None,
                                              (inp["feats"], inp["boxes"]), inp["visn_mask"])
    assert (lang.cpu() - ref["lang"]).abs().max().item() < 2e-2
    assert (visn.cpu() - ref["visn"]).abs().max().item() < 2e-2
    assert (pooled.cpu() - ref["pooled"]).abs().max().item() < 2e-2
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code import _runtime
    _runtime.release()
