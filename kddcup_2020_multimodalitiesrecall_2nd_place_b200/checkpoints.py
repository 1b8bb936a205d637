"""Checkpoint importers: turn what the reference's checkpoints hold into the
{variable name: fp32 array} dict that `mmr_create` / `bind` / `load_state_dict` take.

* TF-1 checkpoints (imagebert_zk, imagebert_lds).  The zk driver restores the EXPONENTIAL-MOVING-AVERAGE shadows:
  `variable_averages.variables_to_restore()` maps every variable `v` to the checkpoint entry
  `v/ExponentialMovingAverage` when it exists (evaluate_normal.py:204-206); the lds driver restores plain names
  through `get_assignment_map_from_checkpoint` and ignores `adam_m` / `adam_v` slots and `global_step`
  (run_pretraining_predict_score.py:343-360).  `tf_variables_from_checkpoint` reads the checkpoint files themselves
  (`<prefix>.index` + `.data-*`, or a directory with a `checkpoint` state file as tf.train.latest_checkpoint resolves
  it, evaluate_normal.py:207-208) through the pure-Python bundle reader of tf_bundle.py -- no TensorFlow needed;
  `tf_variables_from_npz` still takes a `{name: array}` export made elsewhere.
* PyTorch `.pth` state_dicts (lxmert): `KDD.load` (kdd_model.py:131-152) = torch.load + non-strict
  load_state_dict, `module.` prefixes from nn.DataParallel stripped (entry.py:150-158).

No real checkpoint ships with the reference, so these are exercised on synthetic weights: written as a V2 bundle by
tf_bundle.write_bundle (EMA shadows, Adam slots and global_step included), read back, selected, scored.
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping, Tuple

import numpy as np

EMA_SUFFIX = "/ExponentialMovingAverage"
_OPTIMIZER_SLOTS = ("/adam_m", "/adam_v", "/Adam", "/Adam_1", "/Momentum")


def select_tf_variables(entries: Mapping[str, np.ndarray], prefer_ema: bool = True,
                        wanted: Iterable[str] = None) -> Tuple[Dict[str, np.ndarray], Dict[str, str]]:
    """Checkpoint entries -> model variables.  Returns (weights, source) where source[name] is the checkpoint key used.
    prefer_ema=True is the zk driver's behaviour (EMA shadow wins over the raw variable); optimizer slots,
    `global_step` and bookkeeping entries are dropped; a trailing ':0' is tolerated."""
    clean = {(k[:-2] if k.endswith(":0") else k): v for k, v in entries.items()}
    weights, source = {}, {}
    for key in clean:
        if key.endswith(EMA_SUFFIX) or key.endswith(_OPTIMIZER_SLOTS) or key in ("global_step",) \
                or key.startswith(("beta1_power", "beta2_power")):
            continue
        ema = key + EMA_SUFFIX
        use = ema if (prefer_ema and ema in clean) else key
        weights[key] = np.ascontiguousarray(clean[use], dtype=np.float32)
        source[key] = use
    for key in clean:                      # shadows whose raw variable was not saved at all
        if key.endswith(EMA_SUFFIX):
            base = key[: -len(EMA_SUFFIX)]
            if base not in weights and not base.endswith(_OPTIMIZER_SLOTS):
                weights[base] = np.ascontiguousarray(clean[key], dtype=np.float32)
                source[base] = key
    if wanted is not None:
        wanted = list(wanted)
        missing = [w for w in wanted if w not in weights]
        if missing:
            raise KeyError(f"checkpoint lacks {len(missing)} variables, e.g. {missing[:3]}")
        weights = {w: weights[w] for w in wanted}
        source = {w: source[w] for w in wanted}
    return weights, source


def tf_variables_from_npz(path: str, prefer_ema: bool = True, wanted: Iterable[str] = None) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        entries = {k: z[k] for k in z.files}
    return select_tf_variables(entries, prefer_ema=prefer_ema, wanted=wanted)[0]


def tf_variables_from_checkpoint(path: str, prefer_ema: bool = True, wanted: Iterable[str] = None,
                                 verify_data: bool = False) -> Dict[str, np.ndarray]:
    """A TF-1 checkpoint -> weights dict.  `path` is a checkpoint prefix (".../model.ckpt-251") or a directory holding
    a `checkpoint` state file.  zk: prefer_ema=True (evaluate_normal.py:204-212); lds: prefer_ema=False
    (run_pretraining_predict_score.py:347-362)."""
    import os

    from . import tf_bundle
    prefix = path
    if os.path.isdir(path):
        prefix = tf_bundle.latest_checkpoint(path)
        if prefix is None:
            raise FileNotFoundError(f"{path}: no `checkpoint` state file (tf.train.latest_checkpoint would return None)")
    reader = tf_bundle.BundleReader(prefix, verify_data=verify_data)
    return select_tf_variables(reader.read_all(float_only=True), prefer_ema=prefer_ema, wanted=wanted)[0]


def torch_state_dict_to_weights(state_dict: Mapping[str, object]) -> Dict[str, np.ndarray]:
    """`.pth` contents -> weights dict: DataParallel prefix stripped, tensors to fp32 numpy, non-tensor entries
    (e.g. `num_batches_tracked`-style integers) dropped."""
    import torch
    out = {}
    for k, v in state_dict.items():
        name = k[len("module."):] if k.startswith("module.") else k
        if torch.is_tensor(v):
            if not v.is_floating_point():
                continue
            out[name] = v.detach().to(torch.float32).cpu().numpy()
        elif isinstance(v, np.ndarray) and v.dtype.kind == "f":
            out[name] = v.astype(np.float32)
    return out


def load_pth(path: str) -> Dict[str, np.ndarray]:
    """KDD.load (kdd_model.py:131-152): `path` without the `.pth` suffix, as the reference passes it."""
    import torch
    sd = torch.load(path if path.endswith(".pth") else path + ".pth", map_location="cpu")
    return torch_state_dict_to_weights(sd)
