"""CPU-side checks of the reference-named surface: the modules exist under the reference's paths, expose the
reference's names and signatures, and fail loudly (no CPU fallback) when nothing is bound / no GPU is present."""
import inspect

import numpy as np
import pytest
import torch

PKG = "kddcup_2020_multimodalitiesrecall_2nd_place_b200.code"


def test_module_paths_and_signatures_match_the_reference():
    import importlib
    mt = importlib.import_module(PKG + ".imagebert_zk.model_triple")
    assert list(inspect.signature(mt.model_attention_channel_e).parameters) == [
        "num_boxes", "np_boxes_5", "np_images_features", "np_idx_class_labels", "np_len_class_labels",
        "np_idx_query_", "len_query_", "labels", "segment_ids", "label_query", "weight_label_query", "is_training",
        "reuse"]                                                            # model_triple.py:162-163
    pb = importlib.import_module(PKG + ".imagebert_zk.pixelbert")
    assert list(inspect.signature(pb.BertModel.__init__).parameters)[1:10] == [
        "imgfeat", "config", "is_training", "input_ids", "input_mask", "token_type_ids", "use_one_hot_embeddings",
        "scope", "random_sample"]                                           # pixelbert.py:150-158
    pm = importlib.import_module(PKG + ".imagebert_lds.src.pixelmodel")
    assert list(inspect.signature(pm.BertModel.__init__).parameters)[1:11] == [
        "imgfeat", "config", "is_training", "input_ids", "label_ids", "input_mask", "token_type_ids",
        "use_one_hot_embeddings", "scope", "random_sample"]                 # pixelmodel.py:145-154
    rp = importlib.import_module(PKG + ".imagebert_lds.src.run_pretraining_predict_score")
    assert list(inspect.signature(rp.bertmodel).parameters) == [
        "bert_config", "bert_init_checkpoint", "learning_rate", "num_train_steps", "num_warmup_steps",
        "use_one_hot_embeddings", "features", "ngpus", "is_training"]       # run_pretraining_predict_score.py:288
    assert list(inspect.signature(rp.get_next_sentence_output).parameters) == ["bert_config", "input_tensor", "labels"]
    km = importlib.import_module(PKG + ".lxmert.src.tasks.kdd_model")
    assert list(inspect.signature(km.KDDModel.forward).parameters)[1:] == [
        "input_ids", "boxes_label_input_ids", "segment_ids", "input_mask", "boxes_label_segment_ids",
        "boxes_label_input_mask", "feats", "boxes", "visual_attention_mask"]   # kdd_model.py:183-186
    en = importlib.import_module(PKG + ".lxmert.src.lxrt.entry")
    assert list(inspect.signature(en.LXRTEncoder.forward).parameters)[1:] == [
        "input_ids", "boxes_label_input_ids", "segment_ids", "input_mask", "boxes_label_segment_ids",
        "boxes_label_input_mask", "feats", "visual_attention_mask"]            # entry.py:132-135
    mn = importlib.import_module(PKG + ".main")
    assert callable(mn.main)
    cfg = pb.BertConfig.from_dict({"vocab_size": 21128, "hidden_size": 768})
    assert cfg.vocab_size == 21128 and cfg.to_dict()["hidden_size"] == 768


def test_unbound_or_training_calls_fail_loudly():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code import _runtime as rt
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code.imagebert_zk import model_triple
    rt.release()
    z = np.zeros((2, 20), np.int32)
    with pytest.raises(RuntimeError, match="no weights bound"):
        model_triple.model_attention_channel_e(np.ones(2, np.int32), np.zeros((2, 10, 5), np.float32),
                                               np.zeros((2, 10, 2048), np.float32), np.zeros((2, 10, 8), np.int32),
                                               None, z, np.ones(2, np.int32), np.ones(2, np.int64),
                                               np.zeros((2, 30), np.int32), None, None, is_training=False)
    with pytest.raises(NotImplementedError):
        model_triple.model_attention_channel_e(*([None] * 11), is_training=True)
    m = torch.tensor([[1, 1, 0, 1]])
    with pytest.raises(ValueError, match="prefix"):
        rt.prefix_lengths(m, "mask")
    assert rt.prefix_lengths(torch.tensor([[1, 1, 0, 0], [1, 1, 1, 1]]), "mask").tolist() == [2, 4]


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_no_cpu_fallback_behind_the_reference_names():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import synth, _lib
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.code.imagebert_lds.src import run_pretraining_predict_score as rp
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LDS, ModelConfig
    cfg = ModelConfig(LDS, n_layers=1, lq=20, nbox=10, vocab=300)
    w = synth.make_weights(cfg, seed=1)
    inp = synth.make_inputs(cfg, 2, seed=1)
    features = {"input_ids": inp["query_ids"], "segment_ids": inp["segment_ids"], "features": inp["feats"],
                "labelfeat": inp["label_ids"]}
    with pytest.raises((_lib.MmrError, RuntimeError, AssertionError)):
        rp.bertmodel(None, w, 0, 0, 0, False, features, 1, is_training=False)
