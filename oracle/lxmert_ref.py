"""Runs the reference's OWN LXMERT code (unmodified, imported from /root/reference) on CPU.

TEST INFRASTRUCTURE.  Only usable where /root/reference exists (the dev container): it validates
oracle/lxmert.py and generates tests/golden/lxmert_*.npz (tools/make_golden.py).  Nothing that runs on the GPU
box imports this module's reference path.

Staging: param.py parses sys.argv at import (param.py:113) and entry.py hard-codes
"../user_data" (entry.py:116-119), so we pin argv, chdir into a scratch `work/` whose sibling `user_data/` holds
bert_config.json + vocab.txt + an empty pytorch_model.bin (loading is non-strict, modeling.py:805-858).
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile

import torch

REF_SRC = "/root/reference/code/lxmert/src"
REF_USER_DATA = "/root/reference/code/user_data"


def available() -> bool:
    return os.path.isdir(REF_SRC)


def build_reference_model(cfg, weights: dict):
    """Returns the reference KDDModel (eval mode) with `weights` loaded (strict on everything we provide)."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    stage = tempfile.mkdtemp(prefix="lxmert_ref_")
    os.makedirs(os.path.join(stage, "user_data"))
    os.makedirs(os.path.join(stage, "work"))
    bc = json.load(open(os.path.join(REF_USER_DATA, "bert_config.json")))
    bc["vocab_size"] = cfg.vocab
    bc["max_position_embeddings"] = cfg.max_pos
    json.dump(bc, open(os.path.join(stage, "user_data", "bert_config.json"), "w"))
    shutil.copy(os.path.join(REF_USER_DATA, "vocab.txt"), os.path.join(stage, "user_data", "vocab.txt"))
    torch.save({}, os.path.join(stage, "user_data", "pytorch_model.bin"))
    old_argv, old_cwd = sys.argv, os.getcwd()
    rng_state = torch.get_rng_state()
    try:
        sys.argv = ["kdd.py", "--llayers", str(cfg.n_layers), "--rlayers", str(cfg.n_r_layers), "--xlayers",
                    str(cfg.n_x_layers)]
        os.chdir(os.path.join(stage, "work"))
        if REF_SRC not in sys.path:
            sys.path.insert(0, REF_SRC)
        for m in [m for m in sys.modules if m == "param" or m.startswith(("lxrt", "tasks"))]:
            del sys.modules[m]
        import param  # noqa: F401  (parses the pinned argv)
        param.args.load = None
        param.args.llayers, param.args.rlayers, param.args.xlayers = cfg.n_layers, cfg.n_r_layers, cfg.n_x_layers
        from tasks.kdd_model import KDDModel
        model = KDDModel()
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
        torch.set_rng_state(rng_state)
        shutil.rmtree(stage, ignore_errors=True)
    sd = {k: torch.as_tensor(v).clone() for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # everything except the dead MLM head / unused AM-softmax matrix must have been provided
    assert all(k.startswith("cls.") or k == "logit_W" for k in missing), [k for k in missing if not k.startswith("cls.")]
    model.eval()
    return model


@torch.no_grad()
def reference_forward(model, inp):
    """Calls KDDModel.forward exactly as KDD.predict does (kdd_model.py:97-103)."""
    t = lambda a, dt: torch.as_tensor(a).to(dt)  # noqa: E731
    x_norm, lang_scores, logit = model(
        t(inp["query_ids"], torch.long), t(inp["label_ids"], torch.long), None, t(inp["query_mask"], torch.long),
        None, t(inp["label_mask"], torch.long), t(inp["feats"], torch.float32), t(inp["boxes"], torch.float32),
        t(inp["visn_mask"], torch.long))
    score = torch.nn.Softmax(1)(logit)
    return {"x_norm": x_norm, "logit": logit, "probs": score}
